#!/usr/bin/env python3
"""Static issue model of a SASS instruction stream (dev tool; no GPU needed).

Reads `cuobjdump -sass` output, decodes the sm_70+ control field of every instruction (stall count, yield, write/read
barrier, wait mask: bits 105..125 of the 128-bit encoding) and replays a chosen dynamic path through the kernel for W
warps that share one scheduler and one half-rate FP64 pipe (one warp instruction per 2 cycles, measured on B200 by
tools/ubench_fp64.cu).  Output: FP64 instruction count, the single-warp issue time (sum of the stall counts ptxas
wrote = the dependent-issue critical path as the compiler scheduled it), and the modelled FP64 pipe utilisation at W
warps per scheduler.  This is how instruction-level variants of the window sweep are compared before GPU time is spent.

  cuobjdump -sass lib.so > all.sass
  python tools/sass_sched.py all.sass --fun k_solve_tma --contains 'IdLi128' --list          # functions, branches
  python tools/sass_sched.py all.sass --fun ... --path 0x2080-0x4420,0x53f0-0x6390 --warps 2
"""
import argparse
import re
import sys

INS = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/")
ENC2 = re.compile(r"^\s+/\* (0x[0-9a-f]{16}) \*/")

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
VAR_LAT = {"LDS": 33, "LDL": 300, "LDG": 600, "LD": 600, "LDC": 40, "LDCU": 40, "MUFU": 24, "STS": 20, "STL": 20, "STG": 20,
           "SYNCS": 40, "S2R": 30, "S2UR": 30, "ATOMS": 60, "SHFL": 26, "I2F": 20, "F2F": 20, "F2I": 20}


def parse(path):
    funs, cur, name, pend = {}, None, None, None
    for line in open(path):
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            cur = []
            funs[name] = cur
            continue
        if cur is None:
            continue
        m = INS.match(line)
        if m:
            pend = {"addr": int(m.group(1), 16), "text": m.group(2).strip(), "lo": int(m.group(3), 16)}
            continue
        m = ENC2.match(line)
        if m and pend is not None:
            hi = int(m.group(1), 16)
            ctrl = (hi >> 41) & 0x1FFFFF
            pend.update(stall=ctrl & 0xF, yld=(ctrl >> 4) & 1, wbar=(ctrl >> 5) & 7, rbar=(ctrl >> 8) & 7,
                        wait=(ctrl >> 11) & 0x3F)
            t = pend["text"]
            t = re.sub(r"^@!?U?P\w+\s+", "", t)
            pend["op"] = t.split()[0].split(".")[0]
            cur.append(pend)
            pend = None
    return funs


def pick(funs, fun, contains):
    names = [n for n in funs if fun in n and all(c in n for c in contains)]
    if len(names) != 1:
        sys.exit("function match not unique:\n  " + "\n  ".join(names))
    return names[0], funs[names[0]]


def path_ins(ins, spec):
    out = []
    for rng in spec.split(","):
        a, b = (int(x, 16) for x in rng.split("-"))
        out += [i for i in ins if a <= i["addr"] < b]
    return out


def follow(ins, start, end, taken):
    """Dynamic path from address `start` until the instruction at `end` (inclusive): conditional branches are taken iff
    their address is in `taken`, unconditional ones always, CALLs never (the fp64 reciprocal slow path)."""
    by_addr = {i["addr"]: k for k, i in enumerate(ins)}
    k = by_addr[start]
    out = []
    while True:
        i = ins[k]
        out.append(i)
        if i["addr"] == end or len(out) > 100000:
            return out
        if i["op"] == "BRA":
            cond = i["text"].startswith("@")
            if (not cond) or i["addr"] in taken:
                m = re.search(r"(0x[0-9a-f]+)\s*$", i["text"])
                k = by_addr[int(m.group(1), 16)]
                continue
        k += 1


def simulate(seq, warps, reps=3, fp64_cycles=2):
    """W in-order warps, one scheduler (1 issue/cycle), one FP64 pipe busy fp64_cycles per warp instruction."""
    seq = seq * reps
    n = len(seq)
    pc = [0] * warps
    nxt = [w * 7 for w in range(warps)]  # staggered starts
    bars = [[0] * 6 for _ in range(warps)]
    pipe_free = 0
    t = 0
    last = 0
    fp64_issued = 0
    done = 0
    while done < warps:
        issued = False
        for k in range(warps):
            w = (last + 1 + k) % warps
            if pc[w] >= n or nxt[w] > t:
                continue
            i = seq[pc[w]]
            if any((i["wait"] >> b) & 1 and bars[w][b] > t for b in range(6)):
                continue
            is64 = i["op"] in FP64
            if is64 and pipe_free > t:
                continue
            if is64:
                pipe_free = t + fp64_cycles
                fp64_issued += 1
            lat = VAR_LAT.get(i["op"], 20)
            if i["wbar"] != 7:
                bars[w][i["wbar"]] = t + lat
            if i["rbar"] != 7:
                bars[w][i["rbar"]] = max(bars[w][i["rbar"]], t + 10)
            nxt[w] = t + max(1, i["stall"])
            pc[w] += 1
            if pc[w] >= n:
                done += 1
            last = w
            issued = True
            break
        t += 1
        if not issued:
            # jump to the next event
            cand = [nxt[w] for w in range(warps) if pc[w] < n]
            if cand:
                t = max(t, min(min(cand), t + 1))
    return t, fp64_issued


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sass")
    ap.add_argument("--fun", required=True)
    ap.add_argument("--contains", action="append", default=[])
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--path", default=None, help="comma-separated hex address ranges a-b (b exclusive) = the dynamic path")
    ap.add_argument("--follow", default=None, help="start-end (hex): follow the control flow instead of --path")
    ap.add_argument("--taken", default="", help="comma-separated hex addresses of conditional branches that are taken")
    ap.add_argument("--warps", type=int, nargs="*", default=[1, 2, 3, 4])
    a = ap.parse_args()
    funs = parse(a.sass)
    name, ins = pick(funs, a.fun, a.contains)
    print("#", name[:120], len(ins), "instructions")
    if a.list or (a.path is None and a.follow is None):
        for i in ins:
            if i["op"] in ("BRA", "CALL", "RET", "EXIT", "BSSY", "BSYNC", "UTMALDG", "SYNCS", "MUFU", "BAR", "WARPSYNC"):
                print(f"  {i['addr']:#06x}  {i['text'][:90]}")
        return
    if a.follow:
        st, en = (int(x, 16) for x in a.follow.split("-"))
        seq = follow(ins, st, en, {int(x, 16) for x in a.taken.split(",") if x})
    else:
        seq = path_ins(ins, a.path)
    ops = {}
    for i in seq:
        ops[i["op"]] = ops.get(i["op"], 0) + 1
    n64 = sum(v for k, v in ops.items() if k in FP64)
    t1 = sum(max(1, i["stall"]) for i in seq)
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:14]
    print(f"path: {len(seq)} instructions, FP64 {n64} (DFMA {ops.get('DFMA', 0)}, DMUL {ops.get('DMUL', 0)}, DADD {ops.get('DADD', 0)}), "
          f"sum(stall) {t1} cycles, FP64 pipe {2 * n64} cycles -> single-warp pipe use {2 * n64 / t1:.2f}")
    print("      " + ", ".join(f"{k} {v}" for k, v in top))
    for w in a.warps:
        t, f = simulate(seq, w)
        print(f"  {w} warps/scheduler: {t} cycles for {f} FP64 -> FP64 pipe {2 * f / t:.3f}   ({t / (3 * w):.0f} cycles per warp-pass)")


if __name__ == "__main__":
    main()
