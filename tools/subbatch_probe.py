"""Does splitting the 65,536-instance batch of one GPU into independent sub-batches (one handle each, enqueued back to back,
no host sync) recover the 1.73-wave tail of the window-solve kernel?  GPU box:  python tools/subbatch_probe.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from decentralized_ekf_mhe_b200 import build, estimator as E, synth  # noqa: E402

build.build()
N, FILL, K = 20, 26, 200
S = FILL + K


def run(sizes, window_solve=0):
    ests, streams, vos, cs = [], [], [], []
    main = torch.cuda.current_stream()
    for j, n in enumerate(sizes):
        st = {k: v.contiguous() for k, v in synth.make_stream(n, S, seed=100 + j, device="cuda", device_rng=True).items()}
        streams.append(st)
        vos.append([bool(st["vo_flag"][s].any()) for s in range(S)])
        cs.append(torch.cuda.Stream())
        torch.cuda.synchronize()
        with torch.cuda.stream(cs[-1]):  # the handle runs on the stream that is current when it is created
            ests.append(E.BatchedEstimator(E.robot_params("go1", ekf_rate=200, N=N, window_solve=window_solve), n))
    for e, st, vo, c in zip(ests, streams, vos, cs):
        with torch.cuda.stream(c):
            e.run(0, FILL, {k: v[:FILL] for k, v in st.items()}, vo[:FILL])
    torch.cuda.synchronize()
    best = 1e9
    T = FILL
    for rep in range(2):
        k = K // 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(main)
        for e, st, vo, c in zip(ests, streams, vos, cs):
            c.wait_event(e0)
            with torch.cuda.stream(c):
                e.run(T, k, {kk: v[T:T + k] for kk, v in st.items()}, vo[T:T + k])
            d = torch.cuda.Event()
            d.record(c)
            main.wait_event(d)
        e1.record(main)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / k)
        T += k
    for e in ests:
        e.close()
    return best


for ws in (0, 1):
    for sizes in ([65536], [37888, 27648], [32768, 32768], [18944, 18944, 18944, 8704], [16384] * 4):
        ms = run(sizes, ws)
        print(json.dumps({"window_solve": "incremental" if ws else "full", "sub_batches": sizes, "ms_per_tick": ms,
                          "instance_steps_per_s": sum(sizes) / (ms * 1e-3)}), flush=True)
