"""Batch-size sweep helper: runs bench.py for a few --chunk-clips values per model and prints value / ms per step."""
import json
import subprocess
import sys

SWEEPS = {"vggish": (1000, 2000, 5000), "clap-laion-audio": (50, 100, 250), "whisper-small": (32, 64, 128)}

for model, chunks in SWEEPS.items():
    if len(sys.argv) > 1 and model not in sys.argv[1:]:
        continue
    for c in chunks:
        cmd = [sys.executable, "bench.py", "--model", model, "--chunk-clips", str(c), "--steps", "2", "--warmup", "3",
               "--no-cpu-baseline", "--no-e2e"]
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=300).stdout.strip().splitlines()[-1]
            d = json.loads(out)
            print(model, "chunk", c, round(d["value"]), "audio-s/s", round(d["ms_per_step"], 1), "ms", d["clocks"].get("sm_mhz"), flush=True)
        except Exception as e:                                  # noqa: BLE001
            print(model, "chunk", c, "failed:", e, flush=True)
