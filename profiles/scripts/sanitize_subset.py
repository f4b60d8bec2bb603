"""Reduced workload for memory / race checking: every kernel family and every shared-memory or global hand-off the tick
relies on.  Meant for compute-sanitizer (memcheck / racecheck); on pools where the sanitizer is closed it is run against
the -DMPPI_DEBUG_CHECKS=1 build (device-side bounds / protocol assertions trap the kernel) and every handle's guard zones
are verified after each part (mppi_debug_check_guards).
  python profiles/scripts/sanitize_subset.py [part ...]     parts: tick strict racecar tpar batched loop mlp mlp_balanced p2p
Run as: compute-sanitizer --tool memcheck python profiles/scripts/sanitize_subset.py tick strict ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-mppi-mpc_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from golden_util import Golden  # noqa: E402
from gpu_util import engine_from_spec  # noqa: E402
from oracle import mppi_oracle as orc  # noqa: E402

parts = sys.argv[1:] or ["tick", "strict", "racecar", "batched", "loop", "mlp"]
g = Golden("diffdrive_pe0.05")
x0 = np.array([0.1, 0.05, 0.2])


def done(name, arr, eng=None):
    assert np.all(np.isfinite(arr)), name
    if eng is not None:
        assert eng.check_guards() == 0, "guard zone overwritten: " + name
    print("ok", name, "(guards clean)" if eng is not None else "", flush=True)


if "tick" in parts:
    # stash path (noise of the chunk in shared memory, reused as merge scratch by the last CTA), several chunks per CTA and
    # a ragged tail chunk; then the regenerate path (injected noise) and K2 alone from given costs
    sp = orc.diffdrive_spec(K=100000, T=30, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    sp.temperature = 2.0
    eng = engine_from_spec(sp, g.path)
    for t in range(2):
        u0, u = eng.step(x0, None, seed=3, tick=t)
    done("tick stash K=100000 (multi-chunk, ragged tail, 296-partial merge)", u, eng)
    eng.close()
    sp = orc.diffdrive_spec(K=3000, T=30, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    sp.temperature = 2.0
    eng = engine_from_spec(sp, g.path)
    eps = torch.zeros(sp.K, sp.T, 2, device="cuda")
    eng.generate_noise(eps, seed=3, tick=0)
    u0, u = eng.step(x0, eps)
    S = torch.zeros(sp.K, device="cuda")
    eng.rollout_costs(x0, S, eps)
    eng.reduce_update(S, eps)
    done("tick regenerate path (injected noise), K1 alone, K2 alone", u, eng)
    eng.close()
    sp = orc.diffdrive_spec(K=2000, T=61, param_exploration=0.05, cost_mode="last", waypoint_mode="frozen")     # horizon too long to stash
    eng = engine_from_spec(sp, g.path)
    u0, u = eng.step(x0, None, seed=3, tick=0)
    done("tick T=61 (no stash), cost_mode last", u, eng)
    eng.close()

if "strict" in parts:
    sp = orc.diffdrive_spec(K=500, T=30, param_exploration=0.05)
    eng = engine_from_spec(sp, g.path)
    x = x0.copy()
    for t in range(4):
        u0, u = eng.step(x, None, seed=5, tick=t)
        x = orc.plant_diffdrive(x, u0.astype(np.float64), 0.1)
    done("strict multi-pass path (passes %d)" % eng.timings()["last_passes"], u, eng)
    eng.close()

if "racecar" in parts:
    gr = Golden("racecar_default")
    sp = orc.racecar_spec(K=3000, T=50)
    eng = engine_from_spec(sp, gr.path)
    u0, u = eng.step(gr.path[3].astype(np.float64), None, seed=2, tick=0)
    d = torch.zeros(16, 50, 4, device="cuda")
    eng.set_keep_costs(True)
    u0, u = eng.step(gr.path[4].astype(np.float64), None, seed=2, tick=1)
    eng.top_trajectories(gr.path[4].astype(np.float64), d, 16, seed=2, tick=1)
    done("race-car dynamic window + footprint collisions + top-N replay", u, eng)
    eng.close()

if "tpar" in parts or "racecar" in parts:
    # time-parallel rollout (rollout_tpar): ragged CTAs, a single sample, the largest grid it takes, an odd horizon, the
    # graph-captured closed loop; MPPI_TPAR=0 in the environment would send these through the serial rollout instead
    gr = Golden("racecar_default")
    for K, T in ((16384, 50), (37, 50), (1, 7), (296 * 64, 21), (5000, 101)):
        sp = orc.racecar_spec(K=K, T=T)
        eng = engine_from_spec(sp, gr.path)
        for t in range(3):
            u0, u = eng.step(gr.path[3 + t].astype(np.float64), None, seed=2, tick=t)
        done("time-parallel race-car tick K=%d T=%d" % (K, T), u, eng)
        if K == 16384:
            st, ct = eng.run_closed_loop(gr.path[3].astype(np.float64), 6, seed=4, tick0=10, plant=1)
            done("time-parallel closed loop (graph)", st, eng)
        eng.close()

if "batched" in parts or "loop" in parts:
    from mppi_b200.batched import BatchedMPPI
    R = 6
    b = BatchedMPPI(R, g.path, num_samples_K=700, num_horizons_T=20, temperature=2.0, seed=9)
    xs = np.stack([np.append(g.path[10 * r, :2], g.path[10 * r, 2]) for r in range(R)])
    if "batched" in parts:
        u = b.step(torch.from_numpy(xs.astype(np.float32)).cuda().contiguous()).cpu().numpy()
        done("batched fleet tick (one CTA per robot, in-CTA merge)", u, b.engine)
    if "loop" in parts:
        st, ct = b.run_closed_loop(xs, 4)
        st, ct = b.run_closed_loop(st[-1], 4)
        done("fleet closed loop as a CUDA graph (device-side tick counter, grid-wide ticket)", st, b.engine)
    b.engine.close()

if "mlp" in parts or "mlp_balanced" in parts:
    cases = []
    if "mlp" in parts:
        cases += [(2048, 12, 3, 2), (2048, 11, 5, 3)]                    # static schedule: one tile per CTA
    if "mlp_balanced" in parts:
        cases += [(40000, 10, 3, 2), (40000, 11, 5, 3)]                  # ping-pong + balanced hand-off between clusters
    for K, T, n_in, n_hidden in cases:
        mlp = orc.make_mlp(seed=1, out_scale=0.05, n_in=n_in, scalers=(n_in == 5), scaler_gain=1.0, n_hidden=n_hidden)
        sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen", model="diffdrive_mlp", mlp=mlp)
        sp.temperature = 2.0
        eng = engine_from_spec(sp, g.path)
        sc = [mlp[k] for k in ("in_mean", "in_scale", "out_mean", "out_scale")] if n_in == 5 else []
        eng.set_mlp([mlp["W%d" % i] for i in range(n_hidden + 2)], [mlp["b%d" % i] for i in range(n_hidden + 2)], *sc)
        u0, u = eng.step(x0, None, seed=4, tick=0)
        done("learned dynamics K=%d T=%d n_in=%d n_hidden=%d" % (K, T, n_in, n_hidden), u, eng)
        eng.close()

if "p2p" in parts:
    import torch.multiprocessing as mp
    from test_gpu_multi import _free_port, _worker
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1 << 13, 30, 3, out, "p2p")) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=900) for _ in range(2))
    for p in procs:
        p.join(timeout=300)
    assert np.array_equal(got[0], got[1])
    done("fused peer-memory exchange on 2 GPUs", got[0])
print("sanitize subset finished")
