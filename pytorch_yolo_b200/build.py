"""Build libyolo_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m pytorch_yolo_b200.build [--force]

No torch headers are involved: the library is plain CUDA C++ behind ``include/yolo_b200.h``.
The .so is git-ignored but travels to the GPU box with the repo snapshot.

Staleness is decided by CONTENT, not by mtime (a snapshot copy does not keep mtimes): the SHA-256 of every source and
header is stored next to the library (``libyolo_b200.so.srchash``); ``build()`` recompiles the translation units whose
inputs changed (in parallel) and relinks.  ``last_action()`` says what the last call did ("compiled" / "reused").
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import json
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libyolo_b200.so")
STAMP = LIB + ".srchash"
OBJ_DIR = os.path.join(ROOT, "build", "obj")
SOURCES = ["decode.cu", "head.cu", "nms.cu", "postproc.cu", "peer.cu", "targets.cu"]
HEADER = os.path.join(ROOT, "include", "yolo_b200.h")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]

_last_action = "none"


def last_action() -> str:
    return _last_action


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: the yolo_b200 CUDA library cannot be built")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def source_hashes() -> dict:
    """{translation unit: hash of (its text + every header in csrc/ + the public header + the flags)}"""
    shared = hashlib.sha256()
    for h in sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))):
        shared.update(_sha(os.path.join(CSRC, h)).encode())
    shared.update(_sha(HEADER).encode())
    shared.update(" ".join(NVCC_FLAGS).encode())
    out = {}
    for s in SOURCES:
        p = os.path.join(CSRC, s)
        if os.path.isfile(p):
            out[s] = hashlib.sha256((shared.hexdigest() + _sha(p)).encode()).hexdigest()
    return out


def _stamp() -> dict:
    try:
        with open(STAMP) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return {}


def _stale() -> bool:
    return not os.path.isfile(LIB) or _stamp().get("sources") != source_hashes()


def build(force: bool = False, verbose: bool = False) -> str:
    global _last_action
    want = source_hashes()
    if not force and os.path.isfile(LIB) and _stamp().get("sources") == want:
        _last_action = "reused"
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    have = _stamp().get("objects", {})
    logs = {}

    def compile_one(src: str):
        obj = os.path.join(OBJ_DIR, src + ".o")
        log = os.path.join(OBJ_DIR, src + ".log")
        if not force and have.get(src) == want[src] and os.path.isfile(obj) and os.path.isfile(log):
            return src, 0, open(log).read()
        res = subprocess.run([nvcc, *NVCC_FLAGS, "-c", "-o", obj, os.path.join(CSRC, src)], capture_output=True, text=True)
        text = res.stdout + res.stderr
        if res.returncode == 0:
            with open(log, "w") as f:
                f.write(text)
        return src, res.returncode, text

    with cf.ThreadPoolExecutor(max_workers=len(want)) as ex:
        for src, rc, text in ex.map(compile_one, list(want)):
            logs[src] = text
            if verbose or rc != 0:
                sys.stderr.write(text)
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {src}")
    objs = [os.path.join(OBJ_DIR, s + ".o") for s in want]
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB + ".tmp", *objs],
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libyolo_b200.so")
    os.replace(LIB + ".tmp", LIB)
    with open(os.path.join(PKG, "build_ptxas.log"), "w") as f:      # registers / spills / shared memory per kernel
        f.write("".join(logs[s] for s in want))
    with open(STAMP, "w") as f:
        json.dump({"sources": want, "objects": want}, f)
    _last_action = "compiled"
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True), last_action())
