"""Repeatability probe of dekf_run at the benchmark size: one reference run, then R runs compared with it bit for bit (per tick: which
ticks / how many instances differ).  Used to chase the transient one-warp deviation described in DESIGN.md section 10.
    python tools/repeat_probe.py [n=65536] [window_solve=1] [DEKF_VO_COMPACT of the compared runs: 0|1] [runs=7]
    DBG_LOCKSTEP=1: lock-step VO arrival instead of ragged."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from decentralized_ekf_mhe_b200 import build, estimator, synth
build.build()
n, N, K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 20, 130
FILL = 34
S = FILL + K
dev = torch.device("cuda", 0)
RAGGED = os.environ.get("DBG_LOCKSTEP", "0") != "1"
st = synth.make_stream(n, S, device=dev, device_rng=True, vo_jitter=RAGGED)
vo = [bool(st["vo_flag"][s].any()) for s in range(S)]
cut = {k: v for k, v in st.items() if torch.is_tensor(v) and v.shape[0] == S}
res = {}
WS = int(sys.argv[2]) if len(sys.argv) > 2 else 1
MODE = sys.argv[3] if len(sys.argv) > 3 else "0"
TAGS = [("plain", "0")] + [(f"run{r}_compact={MODE}", MODE) for r in range(int(sys.argv[4]) if len(sys.argv) > 4 else 7)]
for tag, env in TAGS:
    os.environ["DEKF_VO_COMPACT"] = env
    est = estimator.BatchedEstimator(estimator.robot_params("go1", ekf_rate=200, N=N, window_solve=WS), n)
    outs = {"quat": torch.empty(S, 4, n, dtype=torch.float64, device=dev), "x": torch.empty(S, 9, n, dtype=torch.float64, device=dev),
            "v_body": torch.empty(S, 3, n, dtype=torch.float64, device=dev), "contact": torch.empty(S, 4, n, dtype=torch.uint8, device=dev),
            "status": torch.empty(S, n, dtype=torch.int32, device=dev)}
    est.run(0, FILL, {k: v[:FILL] for k, v in cut.items()}, vo[:FILL], out={k: v[:FILL] for k, v in outs.items()}, out_per_step=True)
    est.run(FILL, K, {k: v[FILL:] for k, v in cut.items()}, vo[FILL:], out={k: v[FILL:] for k, v in outs.items()}, out_per_step=True)
    torch.cuda.synchronize()
    res[tag] = {k: (v.clone() if tag == "plain" else v) for k, v in outs.items()} if tag == "plain" else None
    if tag != "plain":
        for key in ("quat", "x", "status"):
            d = (outs[key][1:] != res["plain"][key][1:])
            if key != "status":
                d = d.any(dim=1)
            per_tick = d.sum(dim=1)
            bad = torch.nonzero(per_tick).flatten()
            if len(bad):
                print(tag, key, "differing ticks:", (bad[:12] + 1).tolist(), "counts", per_tick[bad[:12]].tolist())
                if key == "x":
                    t = int(bad[0]) + 1
                    ii = torch.nonzero(d[t - 1]).flatten()[:8]
                    i0 = int(ii[0]) // 128 * 128
                    dd = (outs["x"][t] - res["plain"]["x"][t]).abs()
                    print("  per-component maxdiff", [float(v) for v in dd.max(dim=1).values])
                    print("  CTA", i0 // 128, "lanes differing", torch.nonzero(d[t - 1][i0:i0 + 128]).flatten().tolist())
                    print("  CTA flags at t", st["vo_flag"][t, i0:i0 + 128].tolist())
                    print("  CTA status at t", outs["status"][t, i0:i0 + 128].tolist())
                    print("  CTA flags at t-1", st["vo_flag"][t - 1, i0:i0 + 128].tolist())
                    print("  CTA flags at t+1", st["vo_flag"][t + 1, i0:i0 + 128].tolist())
                    print("  v_body maxdiff", float((outs["v_body"][t] - res["plain"]["v_body"][t]).abs().max()))
                    print("  first tick", t, "instances", ii.tolist(), "status", outs["status"][t, ii].tolist(), "plain status", res["plain"]["status"][t, ii].tolist(),
                          "maxdiff", float((outs["x"][t] - res["plain"]["x"][t]).abs().max()), "flags", st["vo_flag"][t, ii].tolist())
        print(tag, "checked")
    est.close()
