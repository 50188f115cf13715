"""Host-side decoders of the C-ABI library under AddressSanitizer / UBSan (tests/harness/host_fuzz.cu): damaged Blosc frames,
random and truncated SAM rows, damaged tensor rows, the row formatter at the edges of the reference text.  Skipped where the sanitizer runtime is
not installed."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_decoders_under_sanitizers(tmp_path):
    exe = str(tmp_path / "host_fuzz")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    build = subprocess.run([nvcc, "-O1", "-g", "-std=c++17", "-Xcompiler", "-fsanitize=address", "-Xcompiler", "-fsanitize=undefined",
                            "-Xcompiler", "-fno-omit-frame-pointer", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                            os.path.join(ROOT, "tests", "harness", "host_fuzz.cu"), "-Xlinker", "-lasan", "-Xlinker", "-lubsan"],
                           capture_output=True, text=True)
    if build.returncode != 0 and ("lasan" in build.stderr or "lubsan" in build.stderr or "sanitize" in build.stderr):
        pytest.skip("sanitizer runtime not available: %s" % build.stderr.strip().splitlines()[-1])
    assert build.returncode == 0, build.stderr
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:protect_shadow_gap=0:abort_on_error=0", UBSAN_OPTIONS="halt_on_error=1")
    run = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=600)
    assert run.returncode == 0, run.stderr[-3000:]
    assert "ERROR: AddressSanitizer" not in run.stderr and "runtime error" not in run.stderr, run.stderr[-3000:]
    out = run.stdout.splitlines()
    assert out[0].startswith("valid frame rc=0 n=205 first=a last=z")
    assert "blosc fuzz" in out[1] and "sam fuzz" in out[2] and "decode fuzz" in out[3] and out[4].startswith("format rc=0")
