"""The pipeline section of bench.py on its own (one GPU): python tools/run_bench_pipeline.py [reps]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "face-recognition-cpp-tensorrt_b200"))
from tools import bench_pipeline as bp  # noqa: E402

print(json.dumps(bp.run_gpu(0, reps=int(sys.argv[1]) if len(sys.argv) > 1 else 20)))
