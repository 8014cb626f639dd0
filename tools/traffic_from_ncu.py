"""profiles/traffic.json from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
launch list of `bench.py --profile`: per bench stage, the measured DRAM bytes per C-ABI call (launches of the last step only).
Usage: python tools/traffic_from_ncu.py launches_traffic.csv [marker] > profiles/traffic.json"""
import collections
import csv
import json
import sys

STAGE = {  # kernel substring -> bench stage (ops.StageTimer span name); calls per step divide the per-step sum
    "roi_gather": "roi_gather", "loftr_layer_tc": "loftr_transformer", "loftr_pack": "loftr_transformer",
    "confidence_reduce": "confidence_reduce", "lsq_": "lsq_align", "extract_knots": "extract_knots",
    "delaunay": "interp_linear", "raster_kernel": "interp_linear", "scaffold": "scale_scaffold", "l1_": "l1_scale_align",
    "compact_valid": "l1_scale_align",
    "upsample_nearest": "upsample_nearest", "bias_act_nhwc": "bias_act", "bias_act_generic": "bias_act",
    "bias_leaky_concat": "leaky_concat", "leaky_concat_nhwc": "leaky_concat", "d2s_bias_act": "d2s_bias_act_concat",
    "conv3x3_head": "conv3x3_head", "decoder_tail": "decoder_tail", "bias_act_pad": "bias_act", "upconv_tc": "upconv",
    "subpixel_conv": "subpixel_conv",
}
CALLS = {"roi_gather": 5, "loftr_transformer": 1, "confidence_reduce": 1, "lsq_align": 1, "extract_knots": 1,
         "interp_linear": 1, "l1_scale_align": 1, "scale_scaffold": 1}

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = collections.OrderedDict()
for r in csv.DictReader(lines):
    key = r["ID"]
    d = rows.setdefault(key, {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1)
    d[r["Metric Name"]] = v * mul
launches = list(rows.values())
marker = sys.argv[2] if len(sys.argv) > 2 else "prologue_flag"
marks = [i for i, l in enumerate(launches) if marker in l["name"]]
if marks:
    launches = launches[marks[-1]:]
per = collections.defaultdict(lambda: [0, 0.0, 0.0])
for l in launches:
    for sub, stage in STAGE.items():
        if sub in l["name"]:
            p = per[stage]
            p[0] += 1
            p[1] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
            p[2] += l.get("gpu__time_duration.sum", 0.0)
            break
out = {}
for stage, (n, b, us) in per.items():
    calls = CALLS.get(stage, n)
    out[stage] = int(b / max(calls, 1))
out["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per C-ABI call, averaged over the calls of one step of "
                "`bench.py --profile` under ncu (kernel launches of one call summed)")
print(json.dumps(out, indent=1))
