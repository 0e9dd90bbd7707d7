# gpufLDA.jl -- what a GPU version of src/fLDA.jl looks like over libtmvb.so (include/tmvb.h, tmvb_flda_*).  The reference has
# none: `@gpu` leaves fLDA / fCTM untouched (macros.jl:274-278).  The struct mirrors fLDA.jl:6-28 with Float32 parameters
# (the convention of gpuLDA.jl) and a handle; with it the `@gpu` macro gets a branch like the LDA one (macros.jl:107-147):
# build a gpufLDA from the CPU model's corpus, copy eta / alpha / kappa / beta / Elogtheta / gamma / tau in, train!, copy back.
# Uses LIBTMVB and tmvb_check from julia/gpuLDA.jl.  gpufCTM is the same pattern over tmvb_fctm_create / tmvb_fctm_upload /
# tmvb_fctm_download on top of julia/gpuCTM.jl (every tmvb_ctm_* call applies to a filtered handle).
#
# UNTESTED IN THIS REPOSITORY (no Julia in the build image); the identical C ABI is exercised by the Python host mirror
# topicmodelsvb.jl_b200/gpu_flda.py and tests/test_flda_gpu.py.

mutable struct gpufLDA <: TopicModel
	K::Int
	M::Int
	V::Int
	N::Vector{Int}
	C::Vector{Int}
	corp::Corpus
	topics::VectorList{Int}
	eta::Float64
	alpha::Vector{Float32}
	kappa::Vector{Float32}
	kappa_old::Vector{Float32}
	beta::Matrix{Float32}
	beta_old::Matrix{Float32}
	Elogtheta::VectorList{Float32}
	Elogtheta_old::VectorList{Float32}
	gamma::VectorList{Float32}
	tau::VectorList{Float32}
	tau_old::VectorList{Float32}
	elbo::Float64
	handle::Ptr{Cvoid}
	hdims::NTuple{3,Int}

	function gpufLDA(corp::Corpus, K::Integer)
		check_corp(corp)
		K > 0 || throw(ArgumentError("number of topics must be a positive integer."))   # fLDA.jl:32
		M, V, U = size(corp)
		N = [length(doc) for doc in corp]
		C = [size(doc) for doc in corp]
		topics = [collect(1:V) for _ in 1:K]
		eta = 0.5                                                       # fLDA.jl:39-51
		alpha = ones(Float32, K)
		kappa = Float32.(rand(Dirichlet(V, 1.0)))
		beta = Float32.(rand(Dirichlet(V, 1.0), K)')
		Elogtheta = [fill(Float32(-Base.MathConstants.eulergamma - digamma(K)), K) for _ in 1:M]
		gamma = [ones(Float32, K) for _ in 1:M]
		tau = [fill(Float32(eta), N[d]) for d in 1:M]
		model = new(K, M, V, N, C, copy(corp), topics, eta, alpha, kappa, copy(kappa), beta, copy(beta), Elogtheta, deepcopy(Elogtheta),
			gamma, tau, deepcopy(tau), 0.0, C_NULL, (0, 0, 0))
		finalizer(m -> (m.handle != C_NULL && ccall((:tmvb_flda_destroy, LIBTMVB), Cint, (Ptr{Cvoid},), m.handle); m.handle = C_NULL), model)
		return model
	end
end

function update_buffer!(model::gpufLDA)
	if model.handle != C_NULL && model.hdims != (model.K, model.M, model.V)
		ccall((:tmvb_flda_destroy, LIBTMVB), Cint, (Ptr{Cvoid},), model.handle)
		model.handle = C_NULL
	end
	if model.handle == C_NULL
		h = Ref{Ptr{Cvoid}}(C_NULL)
		tmvb_check(ccall((:tmvb_flda_create, LIBTMVB), Cint, (Ref{Ptr{Cvoid}}, Int64, Int64, Int64, Cint, Ptr{Cvoid}), h, model.K, model.M, model.V, -1, C_NULL))
		model.handle, model.hdims = h[], (model.K, model.M, model.V)
	end
	terms = vcat([doc.terms for doc in model.corp]...) .- 1
	counts = vcat([doc.counts for doc in model.corp]...)
	N_cumsum = cumsum([0; model.N])
	tmvb_check(ccall((:tmvb_flda_set_corpus, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), model.handle, N_cumsum, terms, counts))
	tmvb_check(ccall((:tmvb_flda_upload, LIBTMVB), Cint,
		(Ptr{Cvoid}, Ref{Cdouble}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, Ref(model.eta), model.alpha, model.kappa, Matrix{Float32}(model.beta), hcat(model.Elogtheta...), hcat(model.gamma...), vcat(model.tau...)))
	nothing
end

function update_host!(model::gpufLDA)
	model.handle == C_NULL && return
	K, M, V = model.K, model.M, model.V
	eta = Ref{Cdouble}(0)
	beta, E, g = Matrix{Float32}(undef, K, V), Matrix{Float32}(undef, K, M), Matrix{Float32}(undef, K, M)
	tau = Vector{Float32}(undef, sum(model.N))
	tmvb_check(ccall((:tmvb_flda_download, LIBTMVB), Cint,
		(Ptr{Cvoid}, Ref{Cdouble}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, eta, model.alpha, model.kappa, beta, E, g, tau))
	beta_old, Eo, tau_old = similar(beta), similar(E), similar(tau)
	tmvb_check(ccall((:tmvb_flda_download_old, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, model.kappa_old, beta_old, Eo, tau_old))
	N_cumsum = cumsum([0; model.N])
	model.eta, model.beta, model.beta_old = eta[], beta, beta_old
	model.Elogtheta, model.Elogtheta_old, model.gamma = [E[:,d] for d in 1:M], [Eo[:,d] for d in 1:M], [g[:,d] for d in 1:M]
	model.tau = [tau[N_cumsum[d]+1:N_cumsum[d+1]] for d in 1:M]
	model.tau_old = [tau_old[N_cumsum[d]+1:N_cumsum[d+1]] for d in 1:M]
	nothing
end

function update_elbo!(model::gpufLDA)                                   # fLDA.jl:105-117
	e = Ref{Cdouble}(0)
	tmvb_check(ccall((:tmvb_flda_elbo, LIBTMVB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), model.handle, e))
	model.elbo = e[]
end

function train!(model::gpufLDA; iter::Integer=150, tol::Real=1.0, niter::Integer=1000, ntol::Real=1/model.K^2, viter::Integer=10, vtol::Real=1/model.K^2, checkelbo::Real=1, printelbo::Bool=true)
	all([tol, ntol, vtol] .>= 0)										|| throw(ArgumentError("tolerance parameters must be nonnegative."))
	all([iter, niter, viter] .>= 0)										|| throw(ArgumentError("iteration parameters must be nonnegative."))
	(isa(checkelbo, Integer) & (checkelbo > 0)) | (checkelbo == Inf)	|| throw(ArgumentError("checkelbo parameter must be a positive integer or Inf."))
	(iter == 0 || viter >= 1)											|| throw(ArgumentError("viter must be at least 1."))
	all([isempty(doc) for doc in model.corp]) ? (iter = 0) : update_buffer!(model)      # fLDA.jl:219
	(checkelbo <= iter) && update_elbo!(model)                                            # fLDA.jl:220

	for k in 1:iter
		# update_phi!/update_tau!/update_gamma!/update_Elogtheta! for _ in 1:viter, update_beta!(model, d), update_kappa!(model, d) (fLDA.jl:223-235)
		tmvb_check(ccall((:tmvb_flda_estep, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Cfloat), model.handle, viter, vtol))
		# update_beta!, update_kappa!, update_alpha!, update_eta! (fLDA.jl:236-239)
		tmvb_check(ccall((:tmvb_flda_mstep, LIBTMVB), Cint, (Ptr{Cvoid}, Int64, Cdouble, Cint, Cdouble), model.handle, model.M, Float64(sum(model.C)), niter, ntol))
		if (checkelbo != Inf) && (k % checkelbo == 0)                                      # check_elbo!, modelutils.jl:574-585
			delta_elbo = -(model.elbo - update_elbo!(model))
			printelbo && println(k, " ∆elbo: ", round(delta_elbo, digits=3))
			delta_elbo < tol && break
		end
	end

	(iter > 0) && update_host!(model)
	if iter > 0
		topics = Matrix{Int32}(undef, model.V, model.K)
		tmvb_check(ccall((:tmvb_flda_topics, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Int32}), model.handle, topics))
		model.topics = [Int.(topics[:,i]) for i in 1:model.K]                               # fLDA.jl:246
	end
	nothing
end
