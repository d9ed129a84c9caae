"""Fuzz the ONNX reader (csrc/onnx_reader.h) and the node-pattern matcher (csrc/graph_plan.h) on the CPU under ASan + UBSan.

    python tools/fuzz_graph_loader.py [seed] [cases per graph and mode]

Builds tools/asan_graph_plan.cc (host only: g++ -fsanitize=address,undefined) and feeds it damaged copies of the four tiny surrogate
graphs: byte-level damage (truncation, overwritten bytes anywhere / in the first 4 KB / in the last 8 KB, deleted and duplicated spans)
and structure-aware damage of valid files (node inputs truncated / renamed / reversed, outputs removed, op types and attributes changed,
nodes deleted or duplicated, initializers resized). Every case must end in "ok" or "rejected", never in a sanitizer report or a crash.
Round 2 found and fixed with it: an out-of-bounds read on a node without outputs (now checked indexing), a varint shift >= 64,
memcpy(nullptr, nullptr, 0) on an empty int64 constant — and, by listing WHICH damaged graphs were still accepted, the missing operand-order
/ Slice / perm / epsilon checks of the matcher (tests/test_graph_plan.py pins those)."""
import copy
import glob
import os
import subprocess
import sys
import sysconfig
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from supertonic_b200 import onnx_lite as ol, surrogate  # noqa: E402

KINDS = ("duration_predictor", "text_encoder", "vector_estimator", "vocoder")
OPS = ["Add", "Mul", "MatMul", "Conv", "Transpose", "Reshape", "LayerNormalization", "Softmax", "Erf", "Div", "Slice", "Concat", "Gather",
       "Unsqueeze", "Identity", "Sub", "Exp", "Clip", "ReduceSum", "BatchNormalization", "Sin", "Cos", "CumSum", "Relu", "Tanh"]


def build(td):
    nl = glob.glob(sysconfig.get_paths()["purelib"] + "/include/cudnn_frontend/thirdparty/nlohmann/json.hpp")[0].rsplit("/nlohmann/", 1)[0]
    exe = os.path.join(td, "asan_graph_plan")
    subprocess.check_call(["g++", "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                           "-I", os.path.join(ROOT, "supertonic_b200", "csrc"), "-I", nl, os.path.join(ROOT, "tools", "asan_graph_plan.cc"), "-o", exe])
    return exe


def byte_damage(raw, rng, mode):
    b = bytearray(raw)
    if mode == 0:
        return bytes(b[:int(rng.integers(0, len(b)))])
    if mode in (1, 2, 4):
        for _ in range(int(rng.integers(1, 8))):
            pos = int(rng.integers(0, len(b))) if mode == 1 else int(rng.integers(0, min(len(b), 4096))) if mode == 2 \
                else len(b) - 1 - int(rng.integers(0, min(len(b), 8192)))
            b[pos] = int(rng.integers(0, 256))
    elif mode == 3:
        p = int(rng.integers(0, len(b)))
        del b[p:min(len(b), p + int(rng.integers(1, 64)))]
    else:
        p = int(rng.integers(0, len(b)))
        b[p:p] = b[p:min(len(b), p + int(rng.integers(1, 200)))]
    return bytes(b)


def structure_damage(base, rng):
    m = copy.deepcopy(base)
    nodes = m.graph.nodes
    for _ in range(int(rng.integers(1, 4))):
        k = int(rng.integers(0, len(nodes)))
        nd, mode = nodes[k], int(rng.integers(0, 9))
        if mode == 0 and nd.inputs:
            nd.inputs = nd.inputs[:int(rng.integers(0, len(nd.inputs)))]
        elif mode == 1:
            nd.outputs = []
        elif mode == 2:
            nd.op_type = str(rng.choice(OPS))
        elif mode == 3 and len(nodes) > 1:
            del nodes[k]
        elif mode == 4 and nd.inputs:
            nd.inputs[int(rng.integers(0, len(nd.inputs)))] = str(rng.choice(["", "nope", "text_ids", "text_mask", "latent_mask"]))
        elif mode == 5 and nd.attrs:
            del nd.attrs[str(rng.choice(list(nd.attrs)))]
        elif mode == 6:
            nodes.insert(k, copy.deepcopy(nd))
        elif mode == 7 and len(nd.inputs) >= 2:
            nd.inputs = list(reversed(nd.inputs))
        elif mode == 8 and m.graph.initializers:
            name = list(m.graph.initializers)[int(rng.integers(0, len(m.graph.initializers)))]
            a = m.graph.initializers[name]
            m.graph.initializers[name] = a.reshape(-1)[:max(1, a.size // 2)].copy() if rng.random() < 0.5 else np.zeros((0,), a.dtype)
    return m


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rng = np.random.default_rng(seed)
    root = surrogate.ensure_assets("tiny")
    ok = rejected = bad = 0
    with tempfile.TemporaryDirectory() as td:
        exe = build(td)
        path = os.path.join(td, "case.onnx")

        def run(kind):
            nonlocal ok, rejected, bad
            r = subprocess.run([exe, path, kind], capture_output=True, text=True, errors="replace")
            if r.returncode != 0 or "runtime error" in r.stderr or "AddressSanitizer" in r.stderr:
                bad += 1
                keep = os.path.join(ROOT, "gpurun_out", f"fuzz_bad_{seed}_{kind}_{bad}.onnx")
                os.makedirs(os.path.dirname(keep), exist_ok=True)
                os.replace(path, keep)
                print("BAD", keep, [ln for ln in r.stderr.splitlines() if "error" in ln.lower()][:2])
            elif r.stdout.startswith("ok"):
                ok += 1
            else:
                rejected += 1
        for kind in KINDS:
            src = os.path.join(root, "onnx", kind + ".onnx")
            raw = open(src, "rb").read()
            base = ol.load_model(src)
            base.metadata = {}
            for mode in range(6):
                for _ in range(n):
                    open(path, "wb").write(byte_damage(raw, rng, mode))
                    run(kind)
            for _ in range(6 * n):
                try:
                    ol.save_model(structure_damage(base, rng), path)
                except Exception:
                    continue
                run(kind)
    print(f"seed {seed}: accepted {ok}, rejected {rejected}, sanitizer reports / crashes {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
