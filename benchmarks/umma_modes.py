"""tcgen05 issue-pattern microbenchmark (csrc/umma_bench.cuh): what one hi/lo-split K step (128 x 128 x 64) costs on the
tensor pipe of every SM at once, operands resident in shared memory - the ways of applying the low weight parts compared
without TMA / epilogue effects.  One JSON line.   python benchmarks/umma_modes.py [--ksteps 200000]"""
import argparse
import json
import os
import sys
from pathlib import Path

os.environ.setdefault("FADTK_SYNTHETIC", "1")
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from fadtk_b200 import _native  # noqa: E402

MODES = {0: "8 x f16 N128 K16 (hi, lo per K slice: WMODE 1)", 1: "4 x f16 N256 K16 (hi | lo stacked along N)",
         2: "4 x f16 N128 + 2 x e4m3 N128 K32, alternating per K step (WMODE 2)",
         3: "same MMAs, kinds grouped over 4 K steps", 4: "4 x e4m3 N128 K32 only", 5: "4 x f16 N128 K16 only (unsplit weights)"}
FLOP = {0: 2 * 2 * 128 * 128 * 64, 1: 2 * 2 * 128 * 128 * 64, 2: 2 * 2 * 128 * 128 * 64, 3: 2 * 2 * 128 * 128 * 64,
        4: 2 * 2 * 128 * 128 * 64, 5: 2 * 128 * 128 * 64}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ksteps", type=int, default=200000)
    ap.add_argument("--modes", default="0,1,2,3,4,5")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    eng = _native.Engine(0, max_examples=16)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    out = {"ksteps_per_sm": args.ksteps, "sms": sms, "modes": []}
    for m, name in ((int(v), MODES[int(v)]) for v in args.modes.split(",")):
        ms = min(eng.umma_mode_ms(m, args.ksteps) for _ in range(3))
        out["modes"].append({"mode": m, "pattern": name, "ms": ms, "ns_per_kstep": ms * 1e6 / args.ksteps,
                             "issued_tflops_chip": FLOP[m] * args.ksteps * sms / (ms * 1e-3) / 1e12})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
