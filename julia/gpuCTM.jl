# gpuCTM.jl -- drop-in replacement for src/gpuCTM.jl of TopicModelsVB.jl: the same `gpuCTM` struct surface (public fields of
# gpuCTM.jl:6-22) and `train!` keywords, every `cl.*` call replaced by a `ccall` into libtmvb.so (include/tmvb.h).  Uses LIBTMVB
# and tmvb_check from julia/gpuLDA.jl.
#
# UNTESTED IN THIS REPOSITORY (Julia is not installed in the build image); the identical C ABI is exercised by the Python host
# mirror topicmodelsvb.jl_b200/gpu_ctm.py and its GPU tests.  To use it: replace `include("gpuCTM.jl")` in
# src/TopicModelsVB.jl by this file and drop the gpuCTM methods of update_buffer!/update_host! (modelutils.jl:400-435,518-537)
# and the `@buffer`/`@host` branches for invsigma / sigma / lambda_dist (macros.jl:67,85,88).

mutable struct gpuCTM <: TopicModel
	K::Int
	M::Int
	V::Int
	N::Vector{Int}
	C::Vector{Int}
	corp::Corpus
	topics::VectorList{Int}
	mu::Vector{Float32}
	sigma::Symmetric{Float32}
	invsigma::Symmetric{Float32}
	beta::Matrix{Float32}
	lambda::VectorList{Float32}
	lambda_dist::Vector{Float32}
	vsq::VectorList{Float32}
	logzeta::Vector{Float32}
	phi::MatrixList{Float32}
	elbo::Float32
	handle::Ptr{Cvoid}          # replaces the 28 OpenCL fields device/context/queue/*_kernel/*_buffer (gpuCTM.jl:23-51)
	hdims::NTuple{3,Int}        # (K, M, V) the handle was created for

	function gpuCTM(corp::Corpus, K::Integer)
		check_corp(corp)
		K > 0 || throw(ArgumentError("number of topics must be a positive integer."))

		M, V, U = size(corp)
		N = [length(doc) for doc in corp]
		C = [size(doc) for doc in corp]
		topics = [collect(1:V) for _ in 1:K]

		mu = zeros(Float32, K)
		sigma = Symmetric(Matrix{Float32}(I, K, K))
		invsigma = copy(sigma)
		beta = rand(Dirichlet(V, 1.0f0), K)'
		lambda = [zeros(Float32, K) for _ in 1:M]
		lambda_dist = zeros(Float32, M)
		vsq = [ones(Float32, K) for _ in 1:M]
		logzeta = fill(0.5f0, M)
		phi = [fill(Float32(1/K), K, N[d]) for d in 1:min(M, 1)]   # materialised on demand (materialize_phi!)
		elbo = 0f0

		model = new(K, M, V, N, C, copy(corp), topics, mu, sigma, invsigma, beta, lambda, lambda_dist, vsq, logzeta, phi, elbo, C_NULL, (0, 0, 0))
		finalizer(m -> (m.handle != C_NULL && ccall((:tmvb_ctm_destroy, LIBTMVB), Cint, (Ptr{Cvoid},), m.handle); m.handle = C_NULL), model)
		return model
	end
end

## update_buffer!(model::gpuCTM)  (modelutils.jl:400-435)
function update_buffer!(model::gpuCTM)
	if model.handle != C_NULL && model.hdims != (model.K, model.M, model.V)   # @gpu overwrites K, M, V (macros.jl:152-160)
		ccall((:tmvb_ctm_destroy, LIBTMVB), Cint, (Ptr{Cvoid},), model.handle)
		model.handle = C_NULL
	end
	if model.handle == C_NULL
		h = Ref{Ptr{Cvoid}}(C_NULL)
		tmvb_check(ccall((:tmvb_ctm_create, LIBTMVB), Cint, (Ref{Ptr{Cvoid}}, Int64, Int64, Int64, Cint, Ptr{Cvoid}),
			h, model.K, model.M, model.V, -1, C_NULL))
		model.handle = h[]
		model.hdims = (model.K, model.M, model.V)
	end
	terms = vcat([doc.terms for doc in model.corp]...) .- 1
	counts = vcat([doc.counts for doc in model.corp]...)
	N_cumsum = cumsum([0; model.N])
	tmvb_check(ccall((:tmvb_ctm_set_corpus, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), model.handle, N_cumsum, terms, counts))
	# sigma is uploaded; the library inverts it in fp64 (the reference's host `inv`, gpuCTM.jl:203-205)
	tmvb_check(ccall((:tmvb_ctm_upload, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, model.mu, Matrix{Float32}(model.sigma), Matrix{Float32}(model.beta), hcat(model.lambda...), hcat(model.vsq...), model.logzeta))
	nothing
end

## update_host!(model::gpuCTM)  (modelutils.jl:518-537); phi stays on the device unless asked for.
function update_host!(model::gpuCTM)
	model.handle == C_NULL && return
	K, M, V = model.K, model.M, model.V
	sigma, invsigma = Matrix{Float32}(undef, K, K), Matrix{Float32}(undef, K, K)
	beta, lambda, vsq = Matrix{Float32}(undef, K, V), Matrix{Float32}(undef, K, M), Matrix{Float32}(undef, K, M)
	model.logzeta = Vector{Float32}(undef, M)
	tmvb_check(ccall((:tmvb_ctm_download, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, model.mu, sigma, invsigma, beta, lambda, vsq, model.logzeta))
	model.sigma, model.invsigma, model.beta = Symmetric(sigma), Symmetric(invsigma), beta
	model.lambda = [lambda[:,d] for d in 1:M]
	model.vsq = [vsq[:,d] for d in 1:M]
	nothing
end

function materialize_phi!(model::gpuCTM)
	phi = Matrix{Float32}(undef, model.K, sum(model.N))
	tmvb_check(ccall((:tmvb_ctm_materialize_phi, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}), model.handle, phi))
	N_cumsum = cumsum([0; model.N])
	model.phi = [phi[:,N_cumsum[d]+1:N_cumsum[d+1]] for d in 1:model.M]
end

## update_elbo! (gpuCTM.jl:135-142) from device-side partials; mode 1 evaluates an arbitrary device state.
function update_elbo!(model::gpuCTM; mode::Integer=0)
	docs, glob = Ref{Cdouble}(0), Ref{Cdouble}(0)
	tmvb_check(ccall((:tmvb_ctm_elbo, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Int64, Ref{Cdouble}, Ref{Cdouble}), model.handle, mode, model.M, docs, glob))
	model.elbo = docs[] + glob[]
end

function check_elbo!(model::gpuCTM, checkelbo::Real, printelbo::Bool, k::Int, tol::Real)
	if k % checkelbo == 0
		delta_elbo = -(model.elbo - update_elbo!(model))
		printelbo && println(k, " ∆elbo: ", round(delta_elbo, digits=3))
		delta_elbo < tol && return true
	end
	false
end

function train!(model::gpuCTM; iter::Integer=150, tol::Real=1.0, niter::Integer=1000, ntol::Real=1/model.K^2, viter::Integer=10, vtol::Real=1/model.K^2, checkelbo::Real=1, printelbo::Bool=true)
	check_model(model)                                                  # gpuCTM.jl:488 (see INTEGRATION.md for the phi rows)
	all([tol, ntol, vtol] .>= 0)										|| throw(ArgumentError("tolerance parameters must be nonnegative."))
	all([iter, niter, viter] .>= 0)										|| throw(ArgumentError("iteration parameters must be nonnegative."))
	(isa(checkelbo, Integer) & (checkelbo > 0)) | (checkelbo == Inf)	|| throw(ArgumentError("checkelbo parameter must be a positive integer or Inf."))
	(iter == 0 || viter >= 1)											|| throw(ArgumentError("viter must be at least 1 (the fused E-step does not keep a stale phi to scatter)."))
	all([isempty(doc) for doc in model.corp]) ? (iter = 0) : update_buffer!(model)
	(checkelbo <= iter) && update_elbo!(model, mode=1)

	for k in 1:iter
		want = (checkelbo != Inf) && (k % checkelbo == 0)
		# update_phi!/update_logzeta!/update_vsq!/update_lambda! for v in 1:viter + the scatter half of update_beta! (gpuCTM.jl:497-507)
		tmvb_check(ccall((:tmvb_ctm_estep, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Cfloat, Cint, Cfloat, Cint), model.handle, niter, ntol, viter, vtol, want))
		# update_beta!, update_sigma! (+ inv), update_mu! (gpuCTM.jl:509-511)
		tmvb_check(ccall((:tmvb_ctm_mstep, LIBTMVB), Cint, (Ptr{Cvoid}, Int64), model.handle, model.M))
		check_elbo!(model, checkelbo, printelbo, k, tol) && break
	end

	(iter > 0) && update_host!(model)
	if iter > 0
		topics = Matrix{Int32}(undef, model.V, model.K)
		tmvb_check(ccall((:tmvb_ctm_topics, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Int32}), model.handle, topics))
		model.topics = [Int.(topics[:,i]) for i in 1:model.K]            # gpuCTM.jl:517
	end
	nothing
end
