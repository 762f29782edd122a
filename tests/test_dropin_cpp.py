"""Drop-in boundary: the reference application's call pattern (src/app.cpp, src/db.cpp) compiles, as C++11,
against the shim headers face-recognition-cpp-tensorrt_b200/cpp/{common,retinaface,arcface,matmul}.h and links libfr_b200.so.
OpenCV's C++ headers are not in this image; tests/cpp/mock_opencv provides a minimal cv::Mat for the test build only."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "face-recognition-cpp-tensorrt_b200"


def _build(tmp_path, built_lib):
    exe = tmp_path / "dropin_usage"
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", f"-I{ROOT / 'tests' / 'cpp' / 'mock_opencv'}", f"-I{PKG / 'cpp'}",
           str(ROOT / "tests" / "cpp" / "dropin_usage.cpp"), "-o", str(exe), f"-L{built_lib.parent}", "-lfr_b200", f"-Wl,-rpath,{built_lib.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_shim_compiles_and_keeps_error_conventions(tmp_path, built_lib):
    exe = _build(tmp_path, built_lib)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "DROPIN OK" in r.stdout


@pytest.mark.gpu
def test_app_call_pattern_runs_on_gpu(tmp_path, built_lib):
    sys.path.insert(0, str(ROOT))
    from tools import synth_weights as sw
    from tools import make_golden_retina as mgr
    from tools import pack_retina as pr
    from tools import pack_weights as pw

    pr.save_retina(tmp_path / "det.frw", sw.retina_state_dict(False, 11, mgr.DET_CLS_SHIFT), False)
    pw.save_arcface(tmp_path / "arc.frw", sw.arcface_state_dict("ir", 7), "ir")  # IR_50: the deployed model
    exe = _build(tmp_path, built_lib)
    r = subprocess.run([str(exe), str(tmp_path / "det.frw"), str(tmp_path / "arc.frw")], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "DROPIN OK:" in r.stdout
