"""Generates tests/golden/prover_k5.json: ONE tiny proof of the oracle's create_proof restatement (oracle/prover_ref.py, pure
Python integers) on the seeded integer-built circuit of tests/test_oracle_prover.py — shape 2 gate-advice + 1 lookup-advice
column, k = 5.  A restatement golden (the reference prover cannot run here): it freezes today's answer so that later rounds
compare the oracle prover — and through tests/test_gpu_prover.py the resident CUDA prover — with a FIXED file.
Run: python tests/golden/make_golden_prover.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import test_oracle_prover as t

K, A, L, SEL, SEED = 5, 2, 1, True, 4242


def proof():
    res = t.run(K, A, L, SEL, SEED)
    return {
        "shape": {"k": K, "gate_advice": A, "lookup_advice": L, "seed": SEED,
                  "note": "instance: tests/test_oracle_prover.int_instance; blinding rows / random polynomial: random.Random(seed + 1); "
                          "SRS: (3 + 5 i) G monomial, (7 + 11 i) G lagrange"},
        "challenges": {k: hex(v) for k, v in res["challenges"].items()},
        "commitments_affine_montgomery": [c.hex() for c in res["commitments"]],
        "evals_montgomery": [[nm, r, bytes(v.tobytes()).hex()] for (nm, r), v in res["evals"].items()],
    }


if __name__ == "__main__":
    json.dump(proof(), open(os.path.join(HERE, "prover_k5.json"), "w"), indent=1)
    print("wrote prover_k5.json")
