"""BASELINE.json configs[4] -- audio-head stress through the fused head (developer entry; the driver-facing arm is
`bench.py --config c5`, which this calls)."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import head_stress  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--H", type=int, default=768)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    print(json.dumps(head_stress(a.steps, 3, a.B, a.H)))
