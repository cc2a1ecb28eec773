"""bench.py's roofline uses an exact operation tally of the committed kernels; this keeps the constants in
sync with the code by re-counting with a counting scalar type (tests/hostsim/flopcount.cpp)."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_flop_tally_matches_bench_constants():
    src = os.path.join(HERE, "hostsim", "flopcount.cpp")
    out = os.path.join(HERE, "hostsim", "_build", "flopcount")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", out, src])
    tally = json.loads(subprocess.check_output([out]).decode())
    import bench
    for k in ("meas_update", "propagate", "propagate_vo", "meas_update_pp", "propagate_pp", "propagate_vo_pp", "ekf_predict", "ekf_correct",
              "ekf_vo_correct", "assemble_go1"):
        assert bench.FLOPS[k] == tally[k], (k, bench.FLOPS[k], tally[k])
    assert tally["assemble_go1_sincos"] == 12  # 3 sincos per leg instead of the generated code's 14 trig calls
    by, fl = bench.algorithmic_work(20, 10.0)["solve"]
    assert by == 2 * 54 * 8 + 21 * 200 + 24 + 96 + 8 and 18000 < fl < 30000
