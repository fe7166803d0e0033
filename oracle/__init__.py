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
`tests/test_oracle_golden.py`.  For the Double64 path the checker is the reference's own committed Double64
log (`tests/golden/sodium_double64.json`, transcribed by `tools/make_golden_dd.py`).

Modules
-------
fem_reference   monomial-coefficient FEM tables exactly as the reference builds them
xc              Slater exchange and PW92 correlation (in-tree builtin formulas)
xc_dd           the same functionals as a T = Double64 run of the reference evaluates them, in mpmath
                (checker of the double-double kernels; Double64 + functional is parity unpinned, see its header)
brent           Optim.jl `Brent()` as called by the reference's line search
scf             `scf_step!` / `groundstate` for ODA + OptimizedAufbau
"""
