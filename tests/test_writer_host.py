"""Host logic of the background column-file writer (marxb200_set_async_writer; marx_b200/csrc/writer.cpp): per-file append order
through a small thread pool, double-buffer hand-over, header patching, error reporting -- checked on CPU with a C++ harness."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n_files,n_batches,n_threads", [(21, 40, 8), (5, 64, 1), (3, 9, 16)])
def test_async_writer_files_equal_sequential_appends(tmp_path, n_files, n_batches, n_threads):
    exe = str(tmp_path / "writer_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "tools", "hostcheck", "writer_check.cpp"),
                           os.path.join(ROOT, "marx_b200", "csrc", "writer.cpp"), "-o", exe])
    d = tmp_path / "out"
    d.mkdir()
    p = subprocess.run([exe, str(d), str(n_files), str(n_batches), str(n_threads)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.startswith("ok unable to open"), p.stdout + p.stderr
    for f in range(n_files):
        raw = (d / ("col%02d.dat" % f)).read_bytes()
        want = np.concatenate([(f * 1000003 + b * 7919 + np.arange(1000 + 37 * b, dtype=np.uint64)).astype(np.uint32) for b in range(n_batches)])
        assert raw[0] == 0x83 and raw[4:5] == b"J"
        assert int.from_bytes(raw[20:24], "big") == len(want)
        assert np.array_equal(np.frombuffer(raw, dtype="<u4", offset=32), want)
