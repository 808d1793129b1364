"""Times the COMPILED host side of the resident prover (include/h2b200_prover.hpp through tests/cpp/prover_mirror_test.cpp)
beside the Python one on the same instance: python tools/bench_cpp_prover.py [k] [A] [L] [reps].  Wall clock per proof,
host buffers in, proof out; the Python number is the same loop around ProverSession.prove."""
import os, sys, subprocess, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import halo2_lib_b200 as h
import bench, test_cpp_mirror as tcm

k = int(sys.argv[1]) if len(sys.argv) > 1 else 19
A = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
n = 1 << k
ctx = h.Context(0)
rng = np.random.default_rng(77)
g = np.array([0xd35d438dc58f0d9d, 0x0a78eb28f5c70b3d, 0x666ea36f7879462c, 0x0e0a77c19a07df2f,
              0xa6ba871b8b1e1b3a, 0x14f1d651eb8e167b, 0xccdd46def0f28c58, 0x1c14ef83340fbe5e], dtype=np.uint64)
def bases(a0, d):
    sc = np.zeros((n, 4), dtype=np.uint64); sc[:, 0] = a0 + d * np.arange(n, dtype=np.uint64)
    return ctx.g1_fixed_base_mul(g, ctx.field_op(1, 5, sc))
bm, bl = bases(3, 5), bases(7, 11)
params = h.ParamsKZG(ctx, k, g=bm, g_lagrange=bl)
inst = h.synthetic_circuit(ctx, k, rng, A=A, L=L)
cs = h.Circuit(ctx, k, inst["fixed"], inst["sigma"], A=A, L=L)
sess = h.ProverSession(ctx, params, cs)
rnd = bench.uniform_residues(rng, n)
v, lk = np.ascontiguousarray(inst["virtual"]), np.ascontiguousarray(inst["lookup"])
prove = lambda: sess.prove(v.ctypes.data, len(v), rnd.ctypes.data, seed=1, break_points=inst["break_points"],
                           lookup_ptr=lk.ctypes.data if len(lk) else 0, n_lookup=len(lk))
sess.blind_log = []
prove()
blind = np.concatenate(sess.blind_log); sess.blind_log = None
prove(); ctx.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    prove()
ctx.synchronize()
print(f"python ProverSession: {(time.perf_counter() - t0) * 1e3 / reps:.3f} ms per proof over {reps} proofs (pageable host buffers)")
d = tempfile.mkdtemp(prefix="h2b_prover_")
w = lambda name, arr: np.ascontiguousarray(arr, dtype=np.uint64).tofile(os.path.join(d, name))
for nm in cs.fixed_names: w("fixed_%s.bin" % nm, inst["fixed"][nm])
for i, sg in enumerate(inst["sigma"]): w("sigma_%d.bin" % i, sg)
w("witness.bin", v); w("breaks.bin", inst["break_points"]); w("lookup.bin", lk); w("random.bin", rnd); w("blind.bin", blind)
w("bases_m.bin", bm); w("bases_l.bin", bl)
open(os.path.join(d, "manifest.txt"), "w").write("%d %d %d 1 %d %d %d %d\n" % (k, A, L, len(v), len(inst["break_points"]), len(lk), len(blind)))
sess.free(); cs.free(); params.close(); ctx.close()
tcm.test_cpp_prover_mirror_compiles_and_links()
out = subprocess.run([os.path.join(ROOT, "build", "prover_mirror_test"), d, "--time", str(reps)], capture_output=True, text=True)
print(out.stdout.strip(), out.stderr.strip())
