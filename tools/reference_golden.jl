# reference_golden.jl -- pins the CPU oracle of this repository to the reference itself.
#
# Needs Julia with TopicModelsVB.jl (v1.x) and JSON.jl installed; NOT runnable in the build image (no Julia), which is why
# the oracle is "parity unpinned" until somebody runs this once and commits the three small JSON files it writes:
#
#     python tools/export_reference_inputs.py          # tests/golden/reference/inputs/*  (already committed)
#     julia tools/reference_golden.jl                  # tests/golden/reference/{lda_cfg0,ctm_cfg,ctpf_cfg,flda_cfg,fctm_cfg}.json
#     python -m pytest tests/test_reference_golden_cpu.py
#
# For every case it loads the corpus with the reference's own `readcorp`, builds the reference's CPU model, INJECTS the initial
# topic table (Julia's RNG stream cannot be reproduced elsewhere; everything else in the constructors is deterministic:
# LDA.jl:34-44, CTM.jl:38-49, CTPF.jl:81-103), then runs the real `train!` one outer iteration at a time
# (`train!(model, iter=1, tol=0, checkelbo=1)`: LDA.jl:161-191, CTM.jl:185-217, CTPF.jl:344-402, fLDA.jl:214-247, fCTM.jl:249-290) recording `model.elbo`
# after each, and finally dumps the trace and the trained parameters.
using TopicModelsVB, JSON, LinearAlgebra

const ROOT = normpath(joinpath(@__DIR__, ".."))
const IN = joinpath(ROOT, "tests", "golden", "reference", "inputs")
const OUT = joinpath(ROOT, "tests", "golden", "reference")

function load_case(case)
	meta = JSON.parsefile(joinpath(IN, case * "_meta.json"))
	corp = readcorp(docfile=joinpath(IN, case * "_docs.txt"), counts=true)
	K, V = meta["K"], meta["V"]
	corp.vocab = Dict{Int,String}(j => string(j) for j in 1:V)
	if haskey(meta, "U")
		corp.users = Dict{Int,String}(u => string(u) for u in 1:meta["U"])
		rc, rd = meta["R_cumsum"], meta["readers"]
		for (d, doc) in enumerate(corp)
			doc.readers = Int[rd[i] for i in rc[d]+1:rc[d+1]]
			doc.ratings = ones(Int, length(doc.readers))
		end
	end
	init = Matrix{Float64}(undef, K, V)
	read!(joinpath(IN, case * "_init.f64"), init)
	return meta, corp, init
end

function build(meta, corp, init, case)
	K, V = meta["K"], meta["V"]
	ctor = Dict("LDA" => LDA, "CTM" => CTM, "CTPF" => CTPF, "fLDA" => fLDA, "fCTM" => fCTM)[meta["model"]]
	model = ctor(corp, K)
	if meta["model"] == "CTPF"
		model.alef = copy(init); model.alef_old = copy(init)
	else
		model.beta = copy(init); model.beta_old = copy(init)
	end
	if meta["model"] in ("fLDA", "fCTM")                  # the injected initial kappa (fLDA.jl:41-42, fCTM.jl:50-51)
		kappa = Vector{Float64}(undef, V)
		read!(joinpath(IN, case * "_kappa.f64"), kappa)
		model.kappa = copy(kappa); model.kappa_old = copy(kappa)
	end
	return model
end

function run_case(case)
	meta, corp, init = load_case(case)
	K = meta["K"]
	model = build(meta, corp, init, case)
	trace = Float64[]
	for k in 1:meta["iter"]
		train!(model, iter=1, tol=0.0, viter=meta["viter"], checkelbo=1, printelbo=false)
		push!(trace, model.elbo)
	end
	# the ELBO of the initial state: a fresh model, train!(iter=0) does not evaluate it, so call update_elbo! directly
	m0 = build(meta, corp, init, case)
	elbo0 = TopicModelsVB.update_elbo!(m0)
	out = Dict{String,Any}("case" => case, "model" => meta["model"], "elbo" => [elbo0; trace],
		"julia" => string(VERSION), "package" => "TopicModelsVB")
	if meta["model"] == "LDA"
		out["alpha"] = model.alpha
		out["beta"] = vec(model.beta)                       # column-major K x V
		out["gamma"] = vcat(model.gamma...)
	elseif meta["model"] == "CTM"
		out["mu"] = model.mu
		out["sigma"] = vec(Matrix(model.sigma))
		out["beta"] = vec(model.beta)
		out["lambda"] = vcat(model.lambda...)
		out["vsq"] = vcat(model.vsq...)
	elseif meta["model"] == "fLDA"
		out["eta"] = model.eta
		out["alpha"] = model.alpha
		out["kappa"] = model.kappa
		out["beta"] = vec(model.beta)
		out["gamma"] = vcat(model.gamma...)
		out["tau"] = vcat(model.tau...)
	elseif meta["model"] == "fCTM"
		out["mu"] = model.mu
		out["sigma"] = vec(Matrix(model.sigma))
		out["kappa"] = model.kappa
		out["beta"] = vec(model.beta)
		out["lambda"] = vcat(model.lambda...)
		out["tau"] = vcat(model.tau...)
	else
		out["alef"] = vec(model.alef)
		out["he"] = vec(model.he)
		out["bet"] = model.bet; out["vav"] = model.vav; out["dalet"] = model.dalet; out["het"] = model.het
		out["gimel"] = vcat(model.gimel...)
		out["zayin"] = vcat(model.zayin...)
	end
	open(joinpath(OUT, case * ".json"), "w") do f
		JSON.print(f, out)
	end
	println(case, ": ELBO ", out["elbo"][1], " -> ", out["elbo"][end])
end

for case in ("lda_cfg0", "ctm_cfg", "ctpf_cfg", "flda_cfg", "fctm_cfg")
	run_case(case)
end
