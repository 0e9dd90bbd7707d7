# gpuLDA.jl -- drop-in replacement for src/gpuLDA.jl of TopicModelsVB.jl: the same `gpuLDA` struct surface and
# `train!` keywords, with every `cl.*` call replaced by a `ccall` into libtmvb.so (include/tmvb.h).
#
# UNTESTED IN THIS REPOSITORY: Julia is not installed in the build image.  The identical C ABI is exercised by
# the Python host mirror (topicmodelsvb.jl_b200/gpu_lda.py) and its GPU tests; this file shows the binding a
# maintainer of the reference would add.  To use it: put libtmvb.so on the loader path (or set
# ENV["TMVB_LIB"]), replace `include("gpuLDA.jl")` in src/TopicModelsVB.jl by this file and drop the
# gpuLDA methods of update_buffer!/update_host! (modelutils.jl:370-397,501-516) and the `@buffer`/`@host`
# branches for alpha / Elogtheta_sum / Elogtheta_dist (macros.jl:64,79,82).

const LIBTMVB = get(ENV, "TMVB_LIB", "libtmvb.so")

struct TmvbStats
	estep_ms::Cdouble
	mstep_ms::Cdouble
	sweeps::Int64
	kernel_launches::Int64
	h2d_bytes::Int64
	d2h_bytes::Int64
end

function tmvb_check(rc::Cint)
	rc == 0 && return nothing
	msg = unsafe_string(ccall((:tmvb_last_error, LIBTMVB), Cstring, ()))
	rc == -5 && throw(TopicModelError(msg))
	rc < 0 && throw(ArgumentError(msg))
	throw(TopicModelError("CUDA error $rc: $msg"))
end

mutable struct gpuLDA <: TopicModel
	K::Int
	M::Int
	V::Int
	N::Vector{Int}
	C::Vector{Int}
	corp::Corpus
	topics::VectorList{Int}
	alpha::Vector{Float32}
	beta::Matrix{Float32}
	Elogtheta::VectorList{Float32}
	Elogtheta_sum::Vector{Float32}
	Elogtheta_dist::Vector{Float32}
	gamma::VectorList{Float32}
	phi::MatrixList{Float32}
	elbo::Float32
	handle::Ptr{Cvoid}          # replaces the 20 OpenCL fields device/context/queue/*_kernel/*_buffer (gpuLDA.jl:21-44)
	hdims::NTuple{3,Int}        # (K, M, V) the handle was created for: update_buffer! recreates it only when they change
	peers::Bool                 # true once connect_peers! has mapped the other ranks' buffers (multi-GPU)
	rank::Int
	world::Int
	allgather::Union{Function,Nothing}   # blob::Vector{UInt8} -> concatenation over ranks (MPI.Allgather, sockets, ...)
	allsum::Union{Function,Nothing}      # x::Float64 -> sum over ranks
	M_total::Int                # corpus-wide document count (== M unless this model holds one shard)

	function gpuLDA(corp::Corpus, K::Integer)
		check_corp(corp)
		K > 0 || throw(ArgumentError("number of topics must be a positive integer."))

		M, V, U = size(corp)
		N = [length(doc) for doc in corp]
		C = [size(doc) for doc in corp]
		topics = [collect(1:V) for _ in 1:K]

		alpha = ones(Float32, K)
		beta = rand(Dirichlet(V, 1.0f0), K)'
		Elogtheta = [fill(Float32(-(eulergamma + digamma(K))), K) for _ in 1:M]
		Elogtheta_sum = sum([Elogtheta; [zeros(Float32, K)]])
		Elogtheta_dist = zeros(Float32, M)
		gamma = [ones(Float32, K) for _ in 1:M]
		phi = [fill(Float32(1/K), K, N[d]) for d in 1:min(M, 1)]   # phi is materialised on demand (materialize_phi!)
		elbo = 0f0

		model = new(K, M, V, N, C, copy(corp), topics, alpha, beta, Elogtheta, Elogtheta_sum, Elogtheta_dist, gamma, phi, elbo, C_NULL, (0, 0, 0), false, 0, 1, nothing, nothing, M)
		finalizer(m -> (m.handle != C_NULL && ccall((:tmvb_lda_destroy, LIBTMVB), Cint, (Ptr{Cvoid},), m.handle); m.handle = C_NULL), model)
		return model
	end
end

## Compute E_q[log(P(theta))] ... Elogqz: unchanged host fallbacks are not needed; the ELBO is assembled on the device.

## update_buffer!(model::gpuLDA)  (modelutils.jl:370-397)
function update_buffer!(model::gpuLDA)
	# The handle (device buffers, mapped peers, captured launch graphs) lives as long as the model; it is rebuilt only when
	# K, M or V were overwritten (@gpu does that, macros.jl:113-121) -- and then the peer handshake is redone as well.
	if model.handle != C_NULL && model.hdims != (model.K, model.M, model.V)
		ccall((:tmvb_lda_destroy, LIBTMVB), Cint, (Ptr{Cvoid},), model.handle)
		model.handle = C_NULL
	end
	fresh = model.handle == C_NULL
	if fresh
		h = Ref{Ptr{Cvoid}}(C_NULL)
		tmvb_check(ccall((:tmvb_lda_create, LIBTMVB), Cint, (Ref{Ptr{Cvoid}}, Int64, Int64, Int64, Cint, Ptr{Cvoid}),
			h, model.K, model.M, model.V, -1, C_NULL))
		model.handle = h[]
		model.hdims = (model.K, model.M, model.V)
	end

	terms = vcat([doc.terms for doc in model.corp]...) .- 1
	counts = vcat([doc.counts for doc in model.corp]...)
	N_cumsum = cumsum([0; model.N])
	tmvb_check(ccall((:tmvb_lda_set_corpus, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
		model.handle, N_cumsum, terms, counts))

	Elogtheta = hcat(model.Elogtheta...)
	gamma = hcat(model.gamma...)
	tmvb_check(ccall((:tmvb_lda_upload, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, model.alpha, Matrix{Float32}(model.beta), Elogtheta, gamma))
	(fresh && model.peers) && map_peers!(model)
	nothing
end

## update_host!(model::gpuLDA)  (modelutils.jl:501-516); phi stays on the device unless asked for.
function update_host!(model::gpuLDA)
	model.handle == C_NULL && return
	beta = Matrix{Float32}(undef, model.K, model.V)
	Elogtheta = Matrix{Float32}(undef, model.K, model.M)
	gamma = Matrix{Float32}(undef, model.K, model.M)
	tmvb_check(ccall((:tmvb_lda_download, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
		model.handle, model.alpha, beta, Elogtheta, gamma))
	model.beta = beta
	model.Elogtheta = [Elogtheta[:,d] for d in 1:model.M]
	model.gamma = [gamma[:,d] for d in 1:model.M]
	Esum = Vector{Float64}(undef, model.K)
	tmvb_check(ccall((:tmvb_lda_get_elogtheta_sum, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float64}), model.handle, Esum))
	model.Elogtheta_sum = Float32.(Esum)
end

## phi for every document (K x N_d), only when the caller wants it (check_model, inspection).
function materialize_phi!(model::gpuLDA)
	phi = Matrix{Float32}(undef, model.K, sum(model.N))
	tmvb_check(ccall((:tmvb_lda_materialize_phi, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Float32}), model.handle, phi))
	N_cumsum = cumsum([0; model.N])
	model.phi = [phi[:,N_cumsum[d]+1:N_cumsum[d+1]] for d in 1:model.M]
end

## Update evidence lower bound from device-side partials (replaces gpuLDA.jl:88-128 + the phi read-back of check_elbo!).
## mode 0: assembled by the last estep!/update_beta!/update_alpha! (already summed over ranks by the exchange kernel);
## mode 1: one pass over this rank's documents for an arbitrary device state -- per-rank, so summed over ranks here.
function update_elbo!(model::gpuLDA; mode::Integer=0)
	docs, glob = Ref{Cdouble}(0), Ref{Cdouble}(0)
	tmvb_check(ccall((:tmvb_lda_elbo, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Int64, Ref{Cdouble}, Ref{Cdouble}),
		model.handle, mode, model.M_total, docs, glob))
	d = docs[]
	(mode != 0 && model.peers && model.allsum !== nothing) && (d = model.allsum(d))
	model.elbo = d + glob[]
	return model.elbo
end

## update_alpha! (gpuLDA.jl:132-154): interior-point Newton in fp64 inside the library.
function update_alpha!(model::gpuLDA, niter::Integer, ntol::Real)
	tmvb_check(ccall((:tmvb_lda_update_alpha, LIBTMVB), Cint, (Ptr{Cvoid}, Int64, Cint, Cdouble, Ptr{Float32}),
		model.handle, model.M_total, niter, ntol, model.alpha))
end

## update_beta! (gpuLDA.jl:201-204)
update_beta!(model::gpuLDA) = tmvb_check(ccall((model.peers ? :tmvb_lda_exchange_mstep : :tmvb_lda_mstep, LIBTMVB), Cint, (Ptr{Cvoid},), model.handle))

## Multi-GPU (one Julia process per device; `model.corp` holds the shard d % world == rank, `model.M_total` the corpus-wide
## document count).  Call once, at any time before train!: `allgather(blob::Vector{UInt8})` returns the world * 512 bytes of
## all ranks in rank order and `allsum(x::Float64)` the sum over ranks (any transport: MPI.Allgather / MPI.Allreduce,
## Distributed, sockets).  The mapping itself happens when the handle exists (update_buffer!) and is redone whenever the
## handle is rebuilt; from then on update_beta! is the fused peer-memory kernel (reduce-scatter of the statistics over
## NVLink + normalise + all-gather).
function connect_peers!(model::gpuLDA, rank::Integer, world::Integer, allgather::Function, allsum::Function)
	model.rank, model.world, model.allgather, model.allsum = rank, world, allgather, allsum
	model.peers = world > 1
	(model.peers && model.handle != C_NULL) && map_peers!(model)
	nothing
end

function map_peers!(model::gpuLDA)
	blob = zeros(UInt8, 512)
	tmvb_check(ccall((:tmvb_lda_comm_export, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), model.handle, blob, 512))
	blobs = model.allgather(blob)::Vector{UInt8}                # world * 512 bytes, rank order; also a barrier over the ranks
	tmvb_check(ccall((:tmvb_lda_comm_connect, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}, Int64), model.handle, model.rank, model.world, blobs, 512))
	nothing
end

## A peer that did not reach the exchange makes the device barriers give up (instead of hanging the GPUs); the reduced
## statistics of that iteration are then incomplete, so the condition is fatal: checked at every ELBO read-back and at update_host!.
function check_peers(model::gpuLDA)
	model.peers || return
	st = Ref{Cint}(0)
	tmvb_check(ccall((:tmvb_lda_comm_status, LIBTMVB), Cint, (Ptr{Cvoid}, Ref{Cint}), model.handle, st))
end

## The folded inner loop: update_phi!/update_gamma!/update_Elogtheta! for v in 1:viter (gpuLDA.jl:356-364).
estep!(model::gpuLDA, viter::Integer, vtol::Real, want_elbo::Bool) =
	tmvb_check(ccall((:tmvb_lda_estep, LIBTMVB), Cint, (Ptr{Cvoid}, Cint, Cfloat, Cint), model.handle, viter, vtol, want_elbo))

function check_elbo!(model::gpuLDA, checkelbo::Real, printelbo::Bool, k::Int, tol::Real)
	if k % checkelbo == 0
		delta_elbo = -(model.elbo - update_elbo!(model))
		check_peers(model)
		printelbo && println(k, " ∆elbo: ", round(delta_elbo, digits=3))
		delta_elbo < tol && return true
	end
	false
end

function train!(model::gpuLDA; iter::Integer=150, tol::Real=1.0, niter::Integer=1000, ntol::Real=1/model.K^2, viter::Integer=10, vtol::Real=1/model.K^2, checkelbo::Real=1, printelbo::Bool=true)
	check_model(model)                                                  # gpuLDA.jl:348 (see INTEGRATION.md for the phi rows of check_model)
	all([tol, ntol, vtol] .>= 0)										|| throw(ArgumentError("tolerance parameters must be nonnegative."))
	all([iter, niter, viter] .>= 0)										|| throw(ArgumentError("iteration parameters must be nonnegative."))
	(isa(checkelbo, Integer) & (checkelbo > 0)) | (checkelbo == Inf)	|| throw(ArgumentError("checkelbo parameter must be a positive integer or Inf."))
	(iter == 0 || viter >= 1)											|| throw(ArgumentError("viter must be at least 1 (the fused E-step does not keep a stale phi to scatter)."))
	all([isempty(doc) for doc in model.corp]) && !model.peers ? (iter = 0) : update_buffer!(model)
	(checkelbo <= iter) && update_elbo!(model, mode=1)

	for k in 1:iter
		estep!(model, viter, vtol, (checkelbo != Inf) && (k % checkelbo == 0))
		update_beta!(model)
		update_alpha!(model, niter, ntol)

		if check_elbo!(model, checkelbo, printelbo, k, tol)
			break
		end
	end

	(iter > 0) && (check_peers(model); update_host!(model))
	if iter > 0
		topics = Matrix{Int32}(undef, model.V, model.K)
		tmvb_check(ccall((:tmvb_lda_topics, LIBTMVB), Cint, (Ptr{Cvoid}, Ptr{Int32}), model.handle, topics))
		model.topics = [Int.(topics[:,i]) for i in 1:model.K]
	end
	nothing
end
