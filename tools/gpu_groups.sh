# lanes per ray on the terrain configs (2: 1080p pitched down, 5: 256 cameras at 720p): bench lines with --group 32 / 16 / 8
mkdir -p gpurun_out
for c in 2 5; do for g in 32 16 8; do
python bench.py --config $c --group $g --steps 4 --warmup 3 --no-cpu-baseline --no-extras 2> gpurun_out/grp_${c}_$g.err | tail -1 > gpurun_out/grp_${c}_$g.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/grp_${c}_$g.json")); print("config $c group $g: value %.1f e2e %.1f p1excl %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_frame"]["exclusive_one_view_in_flight"]["phase1_kernel"]))
except Exception as e:
    print("config $c group $g FAILED", e); print(open("gpurun_out/grp_${c}_$g.err").read()[-800:])
PY
done; done
