"""Inventory of a Supertonic asset directory: for each of the four graphs the I/O signature, the op histogram,
the initializer table (count / bytes / largest), and what the library's node-pattern matcher (csrc/graph_plan.h, stc_derive_arch)
makes of it — the layer plan it would run, or the list of nodes it cannot explain. Host only (no GPU).

usage: python tools/onnx_inventory.py [asset_root | onnx_dir] [--json]
       (no argument: the released assets if they can be found — $SUPERTONIC_ASSETS, assets/, baseline/_ref/assets — else the surrogate set)"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supertonic_b200 import assets, capi, onnx_lite as ol      # noqa: E402

KINDS = ("duration_predictor", "text_encoder", "vector_estimator", "vocoder")


def inventory(onnx_dir: str) -> dict:
    out = {"onnx_dir": onnx_dir, "graphs": {}}
    for kind in KINDS:
        path = os.path.join(onnx_dir, kind + ".onnx")
        if not os.path.exists(path):
            out["graphs"][kind] = {"error": "file not found"}
            continue
        m = ol.load_model(path)
        g = m.graph
        ops = collections.Counter(n.op_type for n in g.nodes)
        inits = sorted(((k, list(v.shape), str(v.dtype), int(v.nbytes)) for k, v in g.initializers.items()), key=lambda t: -t[3])
        rec = {"file_bytes": os.path.getsize(path), "nodes": len(g.nodes), "ops": dict(ops.most_common()),
               "inputs": [[v.name, v.elem_type, list(v.shape)] for v in g.inputs], "outputs": [[v.name, v.elem_type, list(v.shape)] for v in g.outputs],
               "initializers": {"count": len(inits), "bytes": sum(t[3] for t in inits), "largest": inits[:8]},
               "metadata_keys": sorted(m.metadata)}
        try:
            plan = capi.derive_arch(path, kind)
            rec["plan"] = {"layers": [l["type"] for l in plan["layers"]], "summary": dict(collections.Counter(l["type"] for l in plan["layers"])),
                           **{k: v for k, v in plan.items() if k not in ("layers", "t")}}
        except capi.StcError as e:
            rec["plan_error"] = str(e)
        out["graphs"][kind] = rec
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    root = args[0] if args else assets.asset_root("full")[0]
    onnx_dir = root if os.path.exists(os.path.join(root, "vocoder.onnx")) else os.path.join(root, "onnx")
    inv = inventory(onnx_dir)
    if "--json" in sys.argv:
        print(json.dumps(inv, indent=1))
        return
    print(f"# {onnx_dir}")
    for kind, r in inv["graphs"].items():
        if "error" in r:
            print(f"\n## {kind}: {r['error']}")
            continue
        print(f"\n## {kind}: {r['nodes']} nodes, {r['initializers']['count']} initializers ({r['initializers']['bytes'] / 1e6:.1f} MB), file {r['file_bytes'] / 1e6:.1f} MB")
        print("inputs : " + ", ".join(f"{n}{s}" for n, _, s in r["inputs"]))
        print("outputs: " + ", ".join(f"{n}{s}" for n, _, s in r["outputs"]))
        print("ops    : " + ", ".join(f"{k}:{v}" for k, v in r["ops"].items()))
        print("largest: " + ", ".join(f"{n}{s}" for n, s, _, _ in r["initializers"]["largest"][:4]))
        if "plan" in r:
            p = r["plan"]
            print("plan   : " + ", ".join(f"{k} x{v}" for k, v in p["summary"].items()) + "  | " +
                  ", ".join(f"{k}={v}" for k, v in p.items() if k not in ("layers", "summary", "kind", "derived_from")))
        else:
            print("plan   : NOT RECOGNISED — " + r["plan_error"])


if __name__ == "__main__":
    main()
