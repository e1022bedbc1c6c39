cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c32.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > $L
timeout 600 python bench.py --steps 20 2>&1 | tail -1 > gpurun_out/bench_r02_after_warp_noise.json
python - >> $L <<'P'
import json
d=json.loads(open("gpurun_out/bench_r02_after_warp_noise.json").read())
print("value",d["value"],"ms",d["ms_per_step"],"nn",d["config"]["nn_only_ms_per_step"],"sel",d["roofline_tree"]["ms_select"],"bak",d["roofline_tree"]["ms_backup"],"clocks",d["clocks"])
P
cat $L
