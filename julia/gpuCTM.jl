# gpuCTM.jl -- ccall shim replacing the OpenCL half of src/gpuCTM.jl (UNTESTED here: Julia is not installed; the same
# ABI is exercised by topicmodelsvb.jl_b200/gpu_ctm.py).  Uses LIBTMVB / tmvb_check from gpuLDA.jl.  The struct keeps
# its public fields (gpuCTM.jl:6-29); the OpenCL fields collapse into `handle::Ptr{Cvoid}`.

function update_buffer!(model::gpuCTM)
	h = Ref{Ptr{Cvoid}}(C_NULL)
	tmvb_check(ccall((:tmvb_ctm_create, LIBTMVB), Cint, (Ref{Ptr{Cvoid}}, Int64, Int64, Int64, Cint, Ptr{Cvoid}), h, model.K, model.M, model.V, -1, C_NULL))
	model.handle = h[]
	terms = vcat([doc.terms for doc in model.corp]...) .- 1
	counts = vcat([doc.counts for doc in model.corp]...)
	N_cumsum = cumsum([0; model.N])
	tmvb_check(ccall((:tmvb_ctm_set_corpus, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), model.handle, N_cumsum, terms, counts))
	tmvb_check(ccall((:tmvb_ctm_upload, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, model.mu, Matrix{Float32}(model.sigma), Matrix{Float32}(model.beta), hcat(model.lambda...), hcat(model.vsq...), model.logzeta))
end

function update_host!(model::gpuCTM)
	K, M, V = model.K, model.M, model.V
	sigma, invsigma = Matrix{Float32}(undef, K, K), Matrix{Float32}(undef, K, K)
	beta, lambda, vsq = Matrix{Float32}(undef, K, V), Matrix{Float32}(undef, K, M), Matrix{Float32}(undef, K, M)
	model.logzeta = Vector{Float32}(undef, M)
	tmvb_check(ccall((:tmvb_ctm_download, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, model.mu, sigma, invsigma, beta, lambda, vsq, model.logzeta))
	model.sigma, model.invsigma, model.beta = Symmetric(sigma), Symmetric(invsigma), beta
	model.lambda = [lambda[:,d] for d in 1:M]
	model.vsq = [vsq[:,d] for d in 1:M]
end

function update_elbo!(model::gpuCTM; mode::Integer=0)
	docs, glob = Ref{Cdouble}(0), Ref{Cdouble}(0)
	tmvb_check(ccall((:tmvb_ctm_elbo, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Int64, Ref{Cdouble}, Ref{Cdouble}), model.handle, mode, model.M, docs, glob))
	model.elbo = docs[] + glob[]
end

function train!(model::gpuCTM; iter::Integer=150, tol::Real=1.0, niter::Integer=1000, ntol::Real=1/model.K^2, viter::Integer=10, vtol::Real=1/model.K^2, checkelbo::Real=1, printelbo::Bool=true)
	all([tol, ntol, vtol] .>= 0)										|| throw(ArgumentError("tolerance parameters must be nonnegative."))
	all([iter, niter, viter] .>= 0)										|| throw(ArgumentError("iteration parameters must be nonnegative."))
	(isa(checkelbo, Integer) & (checkelbo > 0)) | (checkelbo == Inf)	|| throw(ArgumentError("checkelbo parameter must be a positive integer or Inf."))
	all([isempty(doc) for doc in model.corp]) ? (iter = 0) : update_buffer!(model)
	(checkelbo <= iter) && update_elbo!(model, mode=1)
	for k in 1:iter
		want = (checkelbo != Inf) && (k % checkelbo == 0)
		# update_phi!/update_logzeta!/update_vsq!/update_lambda! for v in 1:viter, then the M-step (gpuCTM.jl:497-511)
		tmvb_check(ccall((:tmvb_ctm_estep, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Cfloat, Cint, Cfloat, Cint), model.handle, niter, ntol, viter, vtol, want))
		tmvb_check(ccall((:tmvb_ctm_mstep, LIBTMVB), Cint, (Ptr{Cvoid}, Int64), model.handle, model.M))
		check_elbo!(model, checkelbo, printelbo, k, tol) && break
	end
	(iter > 0) && update_host!(model)
	model.topics = [reverse(sortperm(vec(model.beta[i,:]))) for i in 1:model.K]   # or tmvb_ctm_topics
	nothing
end
