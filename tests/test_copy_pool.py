"""The context's copy threads (csrc/copy_pool.h, host-only): built with g++ and run under ThreadSanitizer when the
toolchain has it. A hang is a failure (timeout)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "copy_pool_test.cpp")


def _build(tmp_path, flags):
    exe = str(tmp_path / "copy_pool_test")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", *flags, SRC, "-o", exe], capture_output=True, text=True)
    return exe if r.returncode == 0 else None


def test_copy_pool_copies_exactly(tmp_path):
    exe = _build(tmp_path, [])
    assert exe, "g++ could not build the copy pool test"
    r = subprocess.run([exe, "4000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


def test_copy_pool_under_thread_sanitizer(tmp_path):
    exe = _build(tmp_path, ["-fsanitize=thread", "-g"])
    if not exe:
        pytest.skip("no ThreadSanitizer runtime in this toolchain")
    r = subprocess.run([exe, "400"], capture_output=True, text=True, timeout=300)
    if "FATAL: ThreadSanitizer" in r.stderr:  # e.g. an unsupported address-space layout in this container
        pytest.skip("ThreadSanitizer cannot run here: " + r.stderr.strip().splitlines()[0])
    assert r.returncode == 0 and "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-3000:]
