# gpuCTPF.jl -- drop-in replacement for src/gpuCTPF.jl of TopicModelsVB.jl: the same `gpuCTPF` struct surface (public fields of
# gpuCTPF.jl:6-44) and `train!` keywords, every `cl.*` call replaced by a `ccall` into libtmvb.so (include/tmvb.h).  Uses LIBTMVB
# and tmvb_check from julia/gpuLDA.jl.
#
# UNTESTED IN THIS REPOSITORY (Julia is not installed in the build image); the identical C ABI is exercised by the Python host
# mirror topicmodelsvb.jl_b200/gpu_ctpf.py and its GPU tests.  To use it: replace `include("gpuCTPF.jl")` in
# src/TopicModelsVB.jl by this file and drop the gpuCTPF methods of update_buffer!/update_host! (modelutils.jl:438-494,540-570)
# and the `@host model.gimel_buffer` branch (macros.jl:91).

mutable struct gpuCTPF <: TopicModel
	K::Int
	M::Int
	V::Int
	U::Int
	N::Vector{Int}
	C::Vector{Int}
	R::Vector{Int}
	corp::Corpus
	topics::VectorList{Int}
	scores::Matrix{Float32}
	libs::VectorList{Int}
	drecs::VectorList{Int}
	urecs::VectorList{Int}
	a::Float32
	b::Float32
	c::Float32
	d::Float32
	e::Float32
	f::Float32
	g::Float32
	h::Float32
	alef::Matrix{Float32}
	he::Matrix{Float32}
	bet::Vector{Float32}
	vav::Vector{Float32}
	gimel::VectorList{Float32}
	gimel_old::VectorList{Float32}
	zayin::VectorList{Float32}
	dalet::Vector{Float32}
	het::Vector{Float32}
	phi::MatrixList{Float32}
	xi::MatrixList{Float32}
	elbo::Float32
	handle::Ptr{Cvoid}          # replaces the 40-odd OpenCL fields device/context/queue/*_kernel/*_buffer (gpuCTPF.jl:45-74)
	hdims::NTuple{4,Int}        # (K, M, V, U) the handle was created for

	function gpuCTPF(corp::Corpus, K::Integer)
		check_corp(corp)
		K > 0 || throw(ArgumentError("number of topics must be a positive integer."))

		M, V, U = size(corp)
		N = [length(doc) for doc in corp]
		C = [size(doc) for doc in corp]
		R = [length(doc.readers) for doc in corp]

		topics = [collect(1:V) for _ in 1:K]
		scores = zeros(Float32, M, U)

		libs = [Int[] for _ in 1:U]
		for d in 1:M, u in corp[d].readers
			push!(libs[u], d)
		end
		urecs = VectorList{Int}(undef, U)
		for u in 1:U
			ur = trues(M)
			ur[libs[u]] .= false
			urecs[u] = findall(ur)
		end
		drecs = VectorList{Int}(undef, M)
		for d in 1:M
			nr = trues(U)
			nr[corp[d].readers] .= false
			drecs[d] = findall(nr)
		end

		a, b, c, d, e, f, g, h = fill(0.1f0, 8)
		alef = exp.(rand(Dirichlet(V, 1.0f0), K)' .- 0.5f0)
		he = ones(Float32, K, U)
		bet = ones(Float32, K)
		vav = ones(Float32, K)
		gimel = [ones(Float32, K) for _ in 1:M]
		gimel_old = deepcopy(gimel)
		zayin = [ones(Float32, K) for _ in 1:M]
		dalet = ones(Float32, K)
		het = ones(Float32, K)
		phi = [fill(Float32(1/K), K, N[d]) for d in 1:min(M, 1)]     # phi and xi never leave the device (fused E-step)
		xi = [fill(Float32(1/2K), 2K, R[d]) for d in 1:min(M, 1)]
		elbo = 0f0

		model = new(K, M, V, U, N, C, R, copy(corp), topics, scores, libs, drecs, urecs, a, b, c, d, e, f, g, h, alef, he, bet, vav,
			gimel, gimel_old, zayin, dalet, het, phi, xi, elbo, C_NULL, (0, 0, 0, 0))
		finalizer(m -> (m.handle != C_NULL && ccall((:tmvb_ctpf_destroy, LIBTMVB), Cint, (Ptr{Cvoid},), m.handle); m.handle = C_NULL), model)
		return model
	end
end

## update_buffer!(model::gpuCTPF)  (modelutils.jl:438-494)
function update_buffer!(model::gpuCTPF)
	dims = (model.K, model.M, model.V, model.U)
	if model.handle != C_NULL && model.hdims != dims            # @gpu overwrites K, M, V, U (macros.jl:197-206)
		ccall((:tmvb_ctpf_destroy, LIBTMVB), Cint, (Ptr{Cvoid},), model.handle)
		model.handle = C_NULL
	end
	if model.handle == C_NULL
		h = Ref{Ptr{Cvoid}}(C_NULL)
		tmvb_check(ccall((:tmvb_ctpf_create, LIBTMVB), Cint, (Ref{Ptr{Cvoid}}, Int64, Int64, Int64, Int64, Cint, Ptr{Cvoid}),
			h, model.K, model.M, model.V, model.U, -1, C_NULL))
		model.handle = h[]
		model.hdims = dims
	end
	terms = vcat([doc.terms for doc in model.corp]...) .- 1
	counts = vcat([doc.counts for doc in model.corp]...)
	readers = [vcat([doc.readers for doc in model.corp]...) .- 1; 0]     # trailing 0: never a NULL pointer when nobody reads anything
	ratings = [vcat([doc.ratings for doc in model.corp]...); 0]
	tmvb_check(ccall((:tmvb_ctpf_set_corpus, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
		model.handle, cumsum([0; model.N]), terms, counts, cumsum([0; model.R]), readers, ratings))
	hyp = Float64[model.a, model.b, model.c, model.d, model.e, model.f, model.g, model.h]
	tmvb_check(ccall((:tmvb_ctpf_upload, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, hyp, model.alef, model.he, model.bet, model.vav, hcat(model.gimel...), hcat(model.zayin...), model.dalet, model.het))
	nothing
end

## update_host!(model::gpuCTPF)  (modelutils.jl:540-570); phi / xi stay on the device.
function update_host!(model::gpuCTPF)
	model.handle == C_NULL && return
	K, M, V, U = model.K, model.M, model.V, model.U
	alef, he = Matrix{Float32}(undef, K, V), Matrix{Float32}(undef, K, U)
	gimel, zayin = Matrix{Float32}(undef, K, M), Matrix{Float32}(undef, K, M)
	tmvb_check(ccall((:tmvb_ctpf_download, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, alef, he, model.bet, model.vav, gimel, zayin, model.dalet, model.het))
	model.alef, model.he = alef, he
	model.gimel = [gimel[:,d] for d in 1:M]
	model.zayin = [zayin[:,d] for d in 1:M]
	nothing
end

## update_elbo! (gpuCTPF.jl:280-286) from device-side partials; mode 1 evaluates an arbitrary device state.
function update_elbo!(model::gpuCTPF; mode::Integer=0)
	docs, glob = Ref{Cdouble}(0), Ref{Cdouble}(0)
	tmvb_check(ccall((:tmvb_ctpf_elbo, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Int64, Ref{Cdouble}, Ref{Cdouble}), model.handle, mode, model.M, docs, glob))
	model.elbo = docs[] + glob[]
end

function check_elbo!(model::gpuCTPF, checkelbo::Real, printelbo::Bool, k::Int, tol::Real)
	if k % checkelbo == 0
		delta_elbo = -(model.elbo - update_elbo!(model))
		printelbo && println(k, " ∆elbo: ", round(delta_elbo, digits=3))
		delta_elbo < tol && return true
	end
	false
end

## scores / urecs / drecs (gpuCTPF.jl:709-731) on the device: one contraction Eeta' (Etheta + Eepsilon) and two segmented
## rankings with the library / reader entries masked out (tmvb_ctpf_recs; include/tmvb.h).
function update_recs!(model::gpuCTPF)
	M, U = model.M, model.U
	scores = Matrix{Float32}(undef, M, U)
	urecs = Vector{Int32}(undef, M * U - sum(model.R))
	drecs = Vector{Int32}(undef, M * U - sum(model.R))
	uoff, doff = Vector{Int64}(undef, U + 1), Vector{Int64}(undef, M + 1)
	tmvb_check(ccall((:tmvb_ctpf_recs, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Int32}, Ptr{Int64}, Ptr{Int32}, Ptr{Int64}),
		model.handle, scores, urecs, uoff, drecs, doff))
	model.scores = scores
	model.urecs = [Int.(urecs[uoff[u]+1:uoff[u+1]]) for u in 1:U]
	model.drecs = [Int.(drecs[doff[d]+1:doff[d+1]]) for d in 1:M]
	nothing
end

function train!(model::gpuCTPF; iter::Integer=150, tol::Real=1.0, viter::Integer=10, vtol::Real=1/model.K^2, checkelbo::Real=1, printelbo::Bool=true)
	check_model(model)                                                  # gpuCTPF.jl:678 (see INTEGRATION.md for the phi / xi rows)
	all([tol, vtol] .>= 0)												|| throw(ArgumentError("tolerance parameters must be nonnegative."))
	all([iter, viter] .>= 0)											|| throw(ArgumentError("iteration parameters must be nonnegative."))
	(isa(checkelbo, Integer) & (checkelbo > 0)) | (checkelbo == Inf)	|| throw(ArgumentError("checkelbo parameter must be a positive integer or Inf."))
	(iter == 0 || viter >= 1)											|| throw(ArgumentError("viter must be at least 1 (the fused E-step does not keep a stale phi to scatter)."))
	all([isempty(doc) for doc in model.corp]) ? (iter = 0) : update_buffer!(model)
	(checkelbo <= iter) && update_elbo!(model, mode=1)

	for k in 1:iter
		want = (checkelbo != Inf) && (k % checkelbo == 0)
		# update_xi!/update_phi!/update_zayin!/update_gimel! for _ in 1:viter + the scatter halves of update_he!/update_alef! (gpuCTPF.jl:687-697)
		tmvb_check(ccall((:tmvb_ctpf_estep, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Cfloat, Cint), model.handle, viter, vtol, want))
		# update_he!, update_alef!, update_dalet!, update_het!, update_bet!, update_vav! (gpuCTPF.jl:699-704)
		tmvb_check(ccall((:tmvb_ctpf_mstep, LIBTMVB), Cint, (Ptr{Cvoid}, Int64), model.handle, model.M))
		check_elbo!(model, checkelbo, printelbo, k, tol) && break
	end

	(iter > 0) && update_host!(model)
	if iter > 0
		topics = Matrix{Int32}(undef, model.V, model.K)
		tmvb_check(ccall((:tmvb_ctpf_topics, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Int32}), model.handle, topics))
		model.topics = [Int.(topics[:,i]) for i in 1:model.K]            # gpuCTPF.jl:706-707
		update_recs!(model)                                              # gpuCTPF.jl:709-731
	end
	nothing
end
