# gpuCTPF.jl -- ccall shim replacing the OpenCL half of src/gpuCTPF.jl (UNTESTED here: Julia is not installed; the same
# ABI is exercised by topicmodelsvb.jl_b200/gpu_ctpf.py).  Uses LIBTMVB / tmvb_check from gpuLDA.jl.

function update_buffer!(model::gpuCTPF)
	h = Ref{Ptr{Cvoid}}(C_NULL)
	tmvb_check(ccall((:tmvb_ctpf_create, LIBTMVB), Cint, (Ref{Ptr{Cvoid}}, Int64, Int64, Int64, Int64, Cint, Ptr{Cvoid}), h, model.K, model.M, model.V, model.U, -1, C_NULL))
	model.handle = h[]
	terms = vcat([doc.terms for doc in model.corp]...) .- 1
	counts = vcat([doc.counts for doc in model.corp]...)
	readers = [vcat([doc.readers for doc in model.corp]...) .- 1; 0]
	ratings = [vcat([doc.ratings for doc in model.corp]...); 0]
	tmvb_check(ccall((:tmvb_ctpf_set_corpus, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
		model.handle, cumsum([0; model.N]), terms, counts, cumsum([0; model.R]), readers, ratings))
	hyp = Float64[model.a, model.b, model.c, model.d, model.e, model.f, model.g, model.h]
	tmvb_check(ccall((:tmvb_ctpf_upload, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, hyp, model.alef, model.he, model.bet, model.vav, hcat(model.gimel...), hcat(model.zayin...), model.dalet, model.het))
end

function update_host!(model::gpuCTPF)
	K, M, V, U = model.K, model.M, model.V, model.U
	alef, he = Matrix{Float32}(undef, K, V), Matrix{Float32}(undef, K, U)
	gimel, zayin = Matrix{Float32}(undef, K, M), Matrix{Float32}(undef, K, M)
	tmvb_check(ccall((:tmvb_ctpf_download, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, alef, he, model.bet, model.vav, gimel, zayin, model.dalet, model.het))
	model.alef, model.he = alef, he
	model.gimel = [gimel[:,d] for d in 1:M]
	model.zayin = [zayin[:,d] for d in 1:M]
end

function update_elbo!(model::gpuCTPF; mode::Integer=0)
	docs, glob = Ref{Cdouble}(0), Ref{Cdouble}(0)
	tmvb_check(ccall((:tmvb_ctpf_elbo, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Int64, Ref{Cdouble}, Ref{Cdouble}), model.handle, mode, model.M, docs, glob))
	model.elbo = docs[] + glob[]
end

function train!(model::gpuCTPF; iter::Integer=150, tol::Real=1.0, viter::Integer=10, vtol::Real=1/model.K^2, checkelbo::Real=1, printelbo::Bool=true)
	all([tol, vtol] .>= 0)												|| throw(ArgumentError("tolerance parameters must be nonnegative."))
	all([iter, viter] .>= 0)											|| throw(ArgumentError("iteration parameters must be nonnegative."))
	(isa(checkelbo, Integer) & (checkelbo > 0)) | (checkelbo == Inf)	|| throw(ArgumentError("checkelbo parameter must be a positive integer or Inf."))
	all([isempty(doc) for doc in model.corp]) ? (iter = 0) : update_buffer!(model)
	(checkelbo <= iter) && update_elbo!(model, mode=1)
	for k in 1:iter
		want = (checkelbo != Inf) && (k % checkelbo == 0)
		# update_xi!/update_phi!/update_zayin!/update_gimel! for _ in 1:viter, then he, alef, dalet, het, bet, vav (gpuCTPF.jl:687-704)
		tmvb_check(ccall((:tmvb_ctpf_estep, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Cfloat, Cint), model.handle, viter, vtol, want))
		tmvb_check(ccall((:tmvb_ctpf_mstep, LIBTMVB), Cint, (Ptr{Cvoid}, Int64), model.handle, model.M))
		check_elbo!(model, checkelbo, printelbo, k, tol) && break
	end
	(iter > 0) && update_host!(model)
	# topics / scores / drecs / urecs exactly as gpuCTPF.jl:706-731 (host side, unchanged)
	nothing
end
