#!/usr/bin/env bash
# Builds oracle/_ref/ref_host from the reference's own sources WHERE THEY LIE (never copied):
#   /root/reference/cpp/helper.cpp  +  oracle/ref_host_driver.cpp  (ours)  +  oracle/ref_stub (ours).
# The reference's neural path (ONNX Runtime + model files) is NOT buildable here — onnxruntime is
# neither installed nor vendored — so only its pure host functions are compiled.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${SUPERTONIC_REFERENCE:-/root/reference}"
[ -f "$ref/cpp/helper.cpp" ] || { echo "reference sources not present; skipping oracle/_ref" >&2; exit 0; }
nl="$(python - <<'PY'
import glob, sys, sysconfig
c = glob.glob(sysconfig.get_paths()["purelib"] + "/include/cudnn_frontend/thirdparty/nlohmann/json.hpp")
print(c[0].rsplit("/nlohmann/", 1)[0] if c else "")
PY
)"
[ -n "$nl" ] || { echo "nlohmann/json.hpp not found" >&2; exit 1; }
mkdir -p "$here/_ref"
# -O2 without -ffast-math: the float32 length math must keep IEEE semantics.
g++ -std=c++17 -O2 -I "$here/ref_stub" -I "$ref/cpp" -I "$nl" \
    "$ref/cpp/helper.cpp" "$here/ref_host_driver.cpp" -o "$here/_ref/ref_host"
echo "built $here/_ref/ref_host"
# The same reference translation unit against oracle/ref_stub_fake (closed-form stand-ins for the four graphs): the reference's
# _infer / call / batch orchestration runs end to end on the CPU -> tests/golden/pipeline_golden.json (oracle/make_golden.py).
g++ -std=c++17 -O2 -DSTC_FAKE_ORT -I "$here/ref_stub_fake" -I "$ref/cpp" -I "$nl" \
    "$ref/cpp/helper.cpp" "$here/ref_host_driver.cpp" -o "$here/_ref/ref_pipe"
echo "built $here/_ref/ref_pipe"

# The drop-in proof: the UNMODIFIED reference CLI (cpp/example_onnx.cpp + cpp/helper.cpp, compiled where they lie)
# against include/ort_shim/onnxruntime_cxx_api.h, linked to libsupertonic_cuda.so instead of libonnxruntime.
# Run by tests/test_dropin_gpu.py on the B200 box. (-O2, no -ffast-math: see above.)
lib="$here/../supertonic_b200/libsupertonic_cuda.so"
if [ -f "$lib" ]; then
    g++ -std=c++17 -O2 -I "$here/../include/ort_shim" -I "$ref/cpp" -I "$nl" \
        "$ref/cpp/helper.cpp" "$ref/cpp/example_onnx.cpp" -o "$here/_ref/example_onnx_stc" \
        -L "$here/../supertonic_b200" -lsupertonic_cuda -Wl,-rpath,'$ORIGIN/../../supertonic_b200'
    echo "built $here/_ref/example_onnx_stc"
fi
