"""CPU oracle for the batched Gaussian belief-propagation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``phylogaussianbeliefprop.jl_b200/``) may import, link or execute anything in
this directory.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and there
only as the checker / the timed CPU baseline.

It is a restatement (NumPy float64 + a C/OpenMP twin in ``oracle/c``) of the
reference's algorithm -- JuliaPhylo/PhyloGaussianBeliefProp.jl v0.0.1 -- with
every function citing the reference file:line it follows (paths relative to
the reference checkout).  The reference is Julia and Julia is not available in
the build image, so the reference itself cannot be executed here; parity is
PINNED instead by the reference's own golden values (test/*.jl known answers
and the docs' jldoctest outputs), see ``tests/test_oracle_goldens.py`` and
``tests/golden/reference_goldens.json``.

Modules
  network       extended-Newick reader + PhyloNetworks-style preorder (host side
                of the reference, restated only so tests can run without Julia)
  clustergraph  moralize / min-fill / clique tree / Bethe / LTRIP / join-graph,
                spanning-tree schedules   (src/clustergraph.jl)
  models        evolutionary models -> linear-Gaussian factors (src/evomodels/*.jl)
  beliefs       CanonicalBelief, scopeindex, allocatebeliefs, assignfactors!,
                MessageResidual  (src/beliefs.jl)
  bp            marginalize / propagate_belief! / calibrate! / integratebelief! /
                regularizebeliefs_* / free_energy  (src/beliefupdates.jl,
                src/calibration.jl, src/clustergraphbeliefs.jl, src/score.jl)
  densemvn      independent dense multivariate-normal likelihood (the check the
                reference's own tests use in their comments)
  synth         synthetic networks / traits for BASELINE.json configs 2-5
"""
