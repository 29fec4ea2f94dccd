"""Records reference-produced golden vectors: tests/golden/ref_<case>.npz.

Runs the REAL reference (sdevries0/MultiTreeGP on jax + diffrax; sources from $MTGP_REFERENCE,
baseline/_ref or /root/reference) on exactly the cases of tools/make_golden.py and stores what the
reference computes next to the versions it ran on:

    fitness        f32[P]      GeneticProgramming.jit_eval(flat_population, data)   (genetic_programming.py:415)
    traj_fitness   f32[n, B]   Evaluator.evaluate_candidate(...)[-1] of the first n individuals
    xs / us        f32[n, B, K, ...] saved states and controls of the same individuals

    python tools/record_reference_golden.py [--candidates 8] [--threefry-partitionable 1]

This image has no jax / diffrax (and no network), so the script cannot run here: it exits with the
reason.  tests/test_reference_golden.py compares the CPU oracle and the CUDA kernels with every
ref_*.npz that exists and is skipped, with that reason, while none does.  Until a ref_*.npz is
committed the parity of this repository with the reference is "unpinned" (DESIGN.md section 4).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=8, help="individuals whose trajectories are stored")
    ap.add_argument("--threefry-partitionable", type=int, default=1)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()

    import reference_runner as rr
    ok, why = rr.available()
    if not ok:
        print(f"record_reference_golden: the reference cannot run here -- {why}", file=sys.stderr)
        return 3
    from make_golden import golden_cases

    for name, lib, pop, data, env_name, lib_args in golden_cases():
        if name == "_rng":
            continue
        # the case must be the one the oracle-made fixture froze
        frozen = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        assert np.array_equal(frozen["pop"], pop) and np.array_equal(frozen["x0"], data.x0), name
        data.threefry_partitionable = bool(args.threefry_partitionable)
        case = rr.ReferenceCase(lib, *lib_args, pop, data, env_name,
                                threefry_partitionable=bool(args.threefry_partitionable))
        fitness = case.fitness()
        n = min(args.candidates, pop.shape[0])
        cands = [case.candidate(i) for i in range(n)]
        out = dict(fitness=fitness, traj_fitness=np.stack([c["traj_fitness"] for c in cands]),
                   xs=np.stack([c["xs"] for c in cands]), versions=np.array(rr.versions()),
                   threefry_partitionable=np.int32(args.threefry_partitionable), n_candidates=np.int32(n))
        if "us" in cands[0]:
            out["us"] = np.stack([c["us"] for c in cands])
        np.savez_compressed(os.path.join(args.out, "ref_" + name + ".npz"), **out)
        print(f"{name}: recorded {fitness.shape[0]} fitness values with {rr.versions()}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
