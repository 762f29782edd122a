"""Writes tests/golden/synth_rows_kat.json: known-answer bits of the synthetic gallery generator (oracle side).
The device generator (synth_rows_kernel) is compared against the same oracle function on the GPU box."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import search_oracle as so  # noqa: E402

kat = so.synth_rows([0, 12345678901], seed=19)
out = {"seed": 19, "rows": [0, 12345678901], "first2_bits": kat[:, :2].view(np.uint32).tolist()}
(ROOT / "tests" / "golden" / "synth_rows_kat.json").write_text(json.dumps(out, indent=1) + "\n")
print(out)
