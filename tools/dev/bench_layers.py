"""Development: per-layer convolution table of one headline step (bench.py's own profile hooks)."""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--no-secondary", "--no-cpu", "--no-e2e", "--steps", "3", "--warmup", "2"],
                     capture_output=True, text=True).stdout.strip().splitlines()[-1]
d = json.loads(out)
r = d["roofline"]
print("value %.1f Mev/s  step %.2f ms  conv %.2f ms  frac %.3f  chunks %s" % (d["value"], d["ms_per_step"], r["conv_ms_per_step"], r["frac"], d["config"]["pixel_model_chunks"]))
for L in r["layers"]:
    print("  %-12s %6.3f ms  %7.1f TF/s  %.3f" % (L["layer"], L["ms"], L["tflops"], L["frac"]))
