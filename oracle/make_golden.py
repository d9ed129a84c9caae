"""ORACLE tooling — generates tests/golden/host_golden.json by RUNNING the unmodified reference
host code (oracle/_ref/ref_host, built by oracle/build_ref.sh from /root/reference/cpp/helper.cpp).

Run in the build container only (needs /root/reference); the JSON it writes is committed and is
what the CPU test-suite and the GPU box use. Inputs are the reference's own sample strings
(file:line cited per case) plus edge cases for each pure host function.
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from supertonic_b200 import surrogate  # noqa: E402

DEFAULT_EN = ("This morning, I took a walk in the park, and the sound of the birds and the breeze was so "
              "pleasant that I stopped for a long time just to listen.")  # cpp/example_onnx.cpp:17
BATCH_EN = "The sun sets behind the mountains, painting the sky in shades of pink and orange."   # test_all.sh:63
BATCH_KO = "오늘 아침에 공원을 산책했는데, 새소리와 바람 소리가 너무 기분 좋았어요."                 # test_all.sh:64
LONGFORM = ("This is a very long text that will be automatically split into multiple chunks. "
            "The system will process each chunk separately and then concatenate them together with "
            "natural pauses between segments. This ensures that even very long texts can be processed "
            "efficiently while maintaining natural speech flow and avoiding memory issues.")         # test_all.sh:70 (shape)


def longform_from_reference():
    """test_all.sh:70 holds the long-form paragraph; read it where it lies when available."""
    p = "/root/reference/test_all.sh"
    if os.path.exists(p):
        import re
        for line in open(p, encoding="utf-8"):
            m = re.match(r'LONGFORM_TEXT="(.*)"\s*$', line)
            if m:
                return m.group(1)
    return LONGFORM


def cases():
    lf = longform_from_reference()
    c = [
        dict(kind="cfg"),
        dict(kind="text", texts=[DEFAULT_EN], langs=["en"], cite="cpp/example_onnx.cpp:17"),
        dict(kind="text", texts=[BATCH_EN, BATCH_KO], langs=["en", "ko"], cite="test_all.sh:63-64"),
        dict(kind="text", texts=["The Q3 revenue rose 12.5% to $4.2 billion, beating estimates."], langs=["en"]),
        dict(kind="text", texts=["Café déjà vu — “été” à Noël… ça va?"], langs=["fr"], cite="SURVEY.md App. G"),
        dict(kind="text", texts=["no final punct 😀 e.g., a_b [x] @home"], langs=["en"], cite="SURVEY.md App. G"),
        dict(kind="text", texts=["El niño comió piñas, ¿verdad? ¡Sí!", "Ação e coração: não é fácil.",
                                 "Où est l'hôtel? Ça coûte très cher, naïve Zoë."], langs=["es", "pt", "fr"]),
        dict(kind="text", texts=["안녕하세요", "값", "ㄱㄴㄷ 한글 Test"], langs=["ko", "ko", "ko"]),
        dict(kind="text", texts=["  spaces   and\ttabs\n\nnewlines  ", "a ,b .c !d ?e ;f :g 'h"], langs=["en", "en"]),
        dict(kind="text", texts=['She said ""hi"" and \'\'bye\'\' ``ok``', "back\\slash ♥☆♡© → ← | / # `x´"],
             langs=["en", "en"]),
        dict(kind="text", texts=["ends with quote”", "ends with ellipsis…", "ends with bracket)", "x"],
             langs=["en", "en", "en", "en"]),
        dict(kind="text", texts=["i.e., that – or ‑ this — done"], langs=["en"]),
        dict(kind="text", texts=["中文字符 and ÿ and €"], langs=["en"]),
        dict(kind="text", texts=["bad lang"], langs=["de"]),
        dict(kind="text", texts=[""], langs=["en"]),
        dict(kind="text", texts=["invalid \xff\xfe utf8 \xc3"], langs=["en"]),
        dict(kind="chunk", text=lf, max_len=300, cite="test_all.sh:70"),
        dict(kind="chunk", text=lf + "\n\n" + lf + " " + lf, max_len=300),
        dict(kind="chunk", text="Dr. Smith went home. Mr. X stayed! Really? Yes.", max_len=20, cite="SURVEY.md App. G"),
        dict(kind="chunk", text=BATCH_KO + " " + BATCH_KO + " 정말요? 네! 그렇습니다.", max_len=120),
        dict(kind="chunk", text="short", max_len=300),
        dict(kind="chunk", text="", max_len=300),
        dict(kind="chunk", text="   \n\n  \n", max_len=300),
        dict(kind="chunk", text="One.  Two.\tThree.\nFour. " + "x" * 50 + ". tail", max_len=12),
        dict(kind="chunk", text="para one. still one.\n\npara two! and more?\n \n\npara three", max_len=25),
        dict(kind="latent_mask", wav_lengths=[3 * 44100, 9 * 44100 + 17], base_chunk_size=512,
             chunk_compress_factor=6, cite="SURVEY.md App. G"),
        dict(kind="latent_mask", wav_lengths=[1, 3072, 3073, 6144], base_chunk_size=512, chunk_compress_factor=6),
        dict(kind="length_mask", lengths=[0, 1, 5, 3]),
        dict(kind="length_mask", lengths=[2, 4], max_len=7),
        dict(kind="sanitize", text=DEFAULT_EN, max_len=20),
        dict(kind="sanitize", text=BATCH_KO, max_len=10),
        dict(kind="sanitize", text="a/b\\c:d*e?f\"g<h>i|j é😀", max_len=40),
        dict(kind="wav", samples=[0.0, 0.5, -0.5, 1.0, -1.0, 1.5, -1.5, 0.99999, 1e-5, -3.0517578e-05, 0.25],
             sample_rate=44100),
        dict(kind="wav", samples=[], sample_rate=22050),
    ]
    return c + random_cases()


def random_cases():
    """Seeded random inputs for the two text functions (the reference's answers on them widen what the oracle restatement and the
    library's regex-free rewrite are pinned to): sentence / paragraph structure with abbreviations and runs of punctuation for
    chunkText at the chunk sizes call() uses (300; 120 for Korean) and at tiny limits; mixed-script strings with the symbols the
    front-end rewrites or drops for UnicodeProcessor::call."""
    import numpy as np
    rng = np.random.default_rng(2024)
    words = ["the", "river", "Dr.", "Mr.", "Mrs.", "e.g.", "i.e.", "etc.", "U.S.", "3.14", "naïve", "café", "오늘", "아침에", "señor", "ação",
             "x", "A", "walked", "along", "thunder", "Inc.", "Prof.", "vs.", "No.", "St.", "Jr.", "Ph.D.", "a.m."]
    seps = [" ", " ", " ", ", ", ". ", "! ", "? ", "... ", ".\n\n", "\n\n", "\n", "?! ", ".  ", "; ", " — ", ".\" ", "。", "! \n \n"]
    out = []
    for i in range(36):
        k = int(rng.integers(1, 60))
        text = "".join(str(rng.choice(words)) + str(rng.choice(seps)) for _ in range(k))
        if i % 3 == 0:
            text = text.strip()
        out.append(dict(kind="chunk", text=text, max_len=(300, 120, int(rng.integers(1, 40)))[i % 3], cite="seeded random (rng 2024)"))
    alphabet = list("abc XYZ,.!?;:'\"()[]{}_@&%$#-–—…“”‘’´`0123456789éñçãõàüöß한국어오늘") + ["\n", "\t", "  ", "e.g.,", "i.e.,", "😀", "♥", "→", "|", "/"]
    for i in range(12):
        n = int(rng.integers(1, 5))
        texts = ["".join(rng.choice(alphabet, size=int(rng.integers(1, 80)))) for _ in range(n)]
        texts = [t if t.strip() else "x" + t for t in texts]
        out.append(dict(kind="text", texts=texts, langs=[str(rng.choice(["en", "ko", "es", "pt", "fr"])) for _ in range(n)],
                        cite="seeded random (rng 2024)"))
    for i in range(3):              # writeWavFile: clamp, scale by 32767, truncate toward zero (cpp/helper.cpp:985-988)
        x = rng.standard_normal(400) * (0.4, 0.9, 1.3)[i]
        x[::37] = (1.0, -1.0, 0.999985, -0.999985, 3.0517578e-05, -3.0517578e-05, 1.5258789e-05, 0.5, -0.5, 2.0, -2.0)[:len(x[::37])]
        out.append(dict(kind="wav", samples=[float(np.float32(v)) for v in x], sample_rate=(44100, 22050, 16000)[i], cite="seeded random (rng 2024)"))
    for i in range(4):              # getLatentMask / lengthToMask (cpp/helper.cpp:740-770)
        out.append(dict(kind="latent_mask", wav_lengths=[int(v) for v in rng.integers(1, 12 * 44100, size=int(rng.integers(1, 7)))],
                        base_chunk_size=512, chunk_compress_factor=6, cite="seeded random (rng 2024)"))
        out.append(dict(kind="length_mask", lengths=[int(v) for v in rng.integers(0, 40, size=int(rng.integers(1, 9)))]))
    # sampleNoisyLatent (cpp/helper.cpp:424-467): float32 durations -> latent length and mask; durations around multiples of the
    # 3072-sample latent frame (where float32 and integer arithmetic could part) and plain random ones
    cs = 512 * 6
    for i in range(8):
        k = rng.integers(1, 90, size=int(rng.integers(1, 6)))
        d = (k * cs + rng.integers(-2, 3, size=len(k))) / 44100.0
        out.append(dict(kind="noisy_latent", duration=[float(np.float32(v)) for v in d], cite="seeded random (rng 2024), frame boundaries"))
    for i in range(8):
        d = rng.uniform(0.3, 25.0, size=int(rng.integers(1, 7)))
        out.append(dict(kind="noisy_latent", duration=[float(np.float32(v)) for v in d], cite="seeded random (rng 2024)"))
    out.append(dict(kind="noisy_latent", duration=[2.0897958278656006], cite="tests/test_host_oracle.py float32-vs-float64 case"))
    for i in range(6):              # sanitizeFilename (cpp/helper.cpp:1070-1111)
        out.append(dict(kind="sanitize", text="".join(rng.choice(alphabet, size=int(rng.integers(1, 60)))), max_len=int(rng.integers(1, 50)),
                        cite="seeded random (rng 2024)"))
    return out


def pipeline_cases(styles):
    """TextToSpeech::call / batch (cpp/helper.cpp:685-734 over _infer :469-683) run end to end by oracle/_ref/ref_pipe: the unmodified
    reference orchestration over the closed-form stand-ins of oracle/ref_stub_fake/onnxruntime_cxx_api.h."""
    lf = longform_from_reference()
    m1, f1 = styles("M1"), styles("F1")
    return [
        dict(kind="call", text=DEFAULT_EN, lang="en", styles=[m1], total_step=5, speed=1.05, silence_duration=0.3, cite="cpp/example_onnx.cpp:17"),
        dict(kind="call", text=lf, lang="en", styles=[f1], total_step=2, speed=1.0, silence_duration=0.3, cite="test_all.sh:70 (three chunks)"),
        dict(kind="call", text=" ".join([BATCH_KO] * 4) + " 정말요? 네!", lang="ko", styles=[m1], total_step=3, speed=1.2, silence_duration=0.25,
             cite="Korean chunks of 120"),
        dict(kind="call", text="short", lang="en", styles=[m1], total_step=1, speed=0.8, silence_duration=0.0),
        dict(kind="batch", texts=[BATCH_EN, BATCH_KO], langs=["en", "ko"], styles=[m1, f1], total_step=5, speed=1.05, cite="test_all.sh:63-64"),
        dict(kind="batch", texts=["El niño comió piñas, ¿verdad? ¡Sí!", "Ação e coração: não é fácil.", "Où est l'hôtel? Ça coûte très cher.", "x"],
             langs=["es", "pt", "fr", "en"], styles=[m1, f1, f1, m1], total_step=10, speed=0.9),
        dict(kind="batch", texts=[DEFAULT_EN], langs=["en"], styles=[m1], total_step=20, speed=2.0),
        dict(kind="call", text=DEFAULT_EN, lang="en", styles=[m1, f1], total_step=2, speed=1.05, silence_duration=0.3, cite="error: two styles"),
        dict(kind="batch", texts=[DEFAULT_EN, "two"], langs=["en", "en"], styles=[m1], total_step=2, speed=1.05, cite="error: count mismatch"),
        dict(kind="batch", texts=["bad"], langs=["de"], styles=[m1], total_step=2, speed=1.05, cite="error: language"),
    ]


def python_reference_answers(res, assets):
    """The same cases through the UNMODIFIED Python reference (py/helper.py, imported from /root/reference with a fake `onnxruntime`
    module that computes the same stand-ins, oracle/fake_graphs.py): each record gets a `py` object — what TextToSpeech.__call__ /
    batch of the Python port returned and the inputs each InferenceSession.run received."""
    import contextlib
    import importlib.util
    import io
    import numpy as np
    from oracle import fake_graphs
    path = "/root/reference/py/helper.py"
    if not os.path.exists(path):
        return
    cfg = json.load(open(os.path.join(assets, "onnx", "tts.json")))
    chunk = int(cfg["ae"]["base_chunk_size"]) * int(cfg["ttl"]["chunk_compress_factor"])
    trace = []
    saved = sys.modules.get("onnxruntime")
    sys.modules["onnxruntime"] = fake_graphs.fake_onnxruntime_module(trace, chunk)
    try:
        spec = importlib.util.spec_from_file_location("supertonic_reference_py_helper", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        with contextlib.redirect_stdout(io.StringIO()):
            tts = mod.load_text_to_speech(os.path.join(assets, "onnx"), False)
        for r in res:
            c = r["case"]
            del trace[:]
            np.random.seed(0)                       # the port draws its noise from np.random.randn; the stand-ins ignore its values
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    style = mod.load_voice_style(c["styles"])
                    if c["kind"] == "call":
                        wav, dur = tts(c["text"], c["lang"], style, c["total_step"], c["speed"], c["silence_duration"])
                    else:
                        wav, dur = tts.batch(c["texts"], c["langs"], style, c["total_step"], c["speed"])
            except Exception as e:                  # the port asserts / raises ValueError where the C++ throws
                r["py"] = dict(error=f"{type(e).__name__}: {e}")
                continue
            wav = np.asarray(wav, np.float32).reshape(-1)
            r["py"] = dict(wav_len=int(wav.size), wav_sum=float(np.cumsum(wav, dtype=np.float64)[-1]),
                           wav_samples=[float(v) for v in np.concatenate([wav[::1009], wav[-1:]])],
                           duration=[float(v) for v in np.asarray(dur, np.float32).reshape(-1)],
                           trace=[{k: v for k, v in t.items() if k != "_out"} for t in trace])
    finally:
        if saved is None:
            sys.modules.pop("onnxruntime", None)
        else:
            sys.modules["onnxruntime"] = saved


def make_pipeline_golden(assets):
    exe = os.path.join(HERE, "_ref", "ref_pipe")
    cs = pipeline_cases(lambda n: os.path.join(assets, "voice_styles", n + ".json"))
    with tempfile.TemporaryDirectory() as td:
        cj = os.path.join(td, "cases.json")
        with open(cj, "w", encoding="utf-8") as f:
            json.dump(cs, f, ensure_ascii=False)
        raw = subprocess.run([exe, cj, os.path.join(assets, "onnx")], check=True, capture_output=True).stdout
    res = json.loads(raw.decode("utf-8"))
    python_reference_answers(res, assets)
    for r in res:
        r["case"]["styles"] = [os.path.basename(p) for p in r["case"]["styles"]]
    out = os.path.join(ROOT, "tests", "golden", "pipeline_golden.json")
    with open(out, "w", encoding="utf-8") as f:
        json.dump(dict(generator="oracle/make_golden.py", source="unmodified /root/reference/cpp/helper.cpp TextToSpeech::call / batch over "
                       "oracle/ref_stub_fake/onnxruntime_cxx_api.h (closed-form stand-ins for the four graphs)", results=res), f, ensure_ascii=True)
    print(f"wrote {out}: {len(res)} cases")
    for r in res:
        print(r["case"]["kind"], r.get("error") or (r["wav_len"], r["duration"], len(r["trace"])),
              "| py:", (r.get("py") or {}).get("error") or ((r.get("py") or {}).get("wav_len"), (r.get("py") or {}).get("duration")))


def main():
    exe = os.path.join(HERE, "_ref", "ref_host")
    if not os.path.exists(exe):
        subprocess.check_call([os.path.join(HERE, "build_ref.sh")])
    assets = surrogate.ensure_assets("tiny")
    cs = cases()
    cs.append(dict(kind="style", paths=[os.path.join(assets, "voice_styles", n + ".json") for n in ("M1", "F1")]))
    with tempfile.TemporaryDirectory() as td:
        for i, c in enumerate(cs):
            if c["kind"] == "wav":
                c["tmp"] = os.path.join(td, f"w{i}.wav")
        # surrogateescape lets the deliberately invalid UTF-8 case reach the C++ side as raw bytes
        cj = os.path.join(td, "cases.json")
        with open(cj, "wb") as f:
            f.write(json.dumps(cs, ensure_ascii=False).encode("utf-8", "surrogateescape"))
        raw = subprocess.run([exe, cj, os.path.join(assets, "onnx")], check=True, capture_output=True).stdout
    res = json.loads(raw.decode("utf-8", "surrogateescape"))
    for r in res:
        r["case"].pop("tmp", None)
        if r["case"]["kind"] == "style":
            r["case"]["paths"] = [os.path.basename(p) for p in r["case"]["paths"]]
    out = os.path.join(ROOT, "tests", "golden", "host_golden.json")
    with open(out, "w", encoding="utf-8", errors="surrogateescape") as f:
        json.dump(dict(generator="oracle/make_golden.py", source="unmodified /root/reference/cpp/helper.cpp",
                       indexer="supertonic_b200.surrogate.build_indexer()", results=res), f, ensure_ascii=True)
    print(f"wrote {out}: {len(res)} cases")
    make_pipeline_golden(assets)
    for r in res:
        k = r["case"]["kind"]
        if "error" in r:
            print(k, "ERROR", r["error"])
        elif k == "text":
            print(k, [len(x) for x in r["text_ids"]], [int(sum(m[0])) for m in r["text_mask"]])
        elif k == "chunk":
            print(k, [len(x.encode("utf-8", "surrogateescape")) for x in r["chunks"]])
        elif k == "noisy_latent":
            print(k, r["latent_shape"], len(r["mask"][0][0]), r["masked_zero"], r["live_nonzero"])


if __name__ == "__main__":
    main()
