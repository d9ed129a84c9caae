for m in 0 2 3; do
  STC_PDL=$m timeout 300 python bench.py --no-strong --no-vary --no-cpu-baseline --steps 10 > gpurun_out/r2y_pdl$m.json 2> gpurun_out/r2y_pdl$m.err; echo "mode $m rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2y_pdl$m.json") if l.startswith("{")][-1])
print("mode $m", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "lat", round(d["latency"]["p50_ms"],3), "long", round(d["latency"]["long_form"]["call_batched_p50_ms"],2), d["parity_check"]["latent_max_abs"], d["parity_check"]["snr_db"], d["roofline"]["in_graph"]["us_per_block_latent_rows"])
PY
done
