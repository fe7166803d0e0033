"""CPU oracle: a restatement of AtomicKohnSham.jl's SCF hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `atomickohnsham_b200/` imports this
package; only `tests/`, `__graft_entry__.smoke()` and the CPU legs of
`bench.py` do, and there only as the checker / the reported CPU baseline.

The reference is pure Julia and neither Julia nor Libxc exists in this image
(SURVEY.md section 0), so the reference itself cannot be executed here.  The
oracle is pinned instead against every golden value the reference commits for
this path: the FEM matrices fixture
(`test/reference_data/fem_matrices.jld2`), the hydrogen and scandium test
energies, and the sodium / oxygen trajectories under `examples/`.  See
`tests/test_oracle_golden.py`.

Modules
-------
fem_reference   monomial-coefficient FEM tables exactly as the reference builds them
xc              Slater exchange and PW92 correlation (in-tree builtin formulas)
brent           Optim.jl `Brent()` as called by the reference's line search
scf             `scf_step!` / `groundstate` for ODA + OptimizedAufbau
"""
