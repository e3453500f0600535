"""profiles/r02_traffic.json from the raw ncu launch lists committed next to it (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per
kernel, `--cache-control none`: L2 is NOT flushed between kernels, as inside the CUDA-graph replay the bench times).  bench.py reads the JSON for
`roofline.traffic` (per launch of the dominant launch = one UNet sample-forward = the sum over its kernels) and `dominant_kernel.traffic`.
usage: python tools/summarize_traffic.py profiles/r02_unet_forward_launches_8samples.csv [profiles/r02_unet_forward_launches_edit_3samples.csv] [profiles/r02_decode_launches_b1.csv]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = r["ID"]
        d = per.setdefault(k, {"name": re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"])), "grid": r.get("Grid Size", "")})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        elif m.startswith("dram__bytes"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d[m.split(".")[0]] = v * scale
    return list(per.values())


def total_bytes(rs):
    return sum(r.get("dram__bytes_read", 0) + r.get("dram__bytes_write", 0) for r in rs)


out = {"source": "profiles/" + ", profiles/".join(os.path.basename(p) for p in sys.argv[1:]) + " (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                 "dram__bytes_write.sum --cache-control none --clock-control none; tools/summarize_traffic.py)"}
for path in sys.argv[1:]:
    rs = rows(path)
    base = os.path.basename(path)
    m = re.search(r"(\d+)samples", base)
    if "unet_forward" in base and m:
        n = int(m.group(1))
        out[f"unet_forward_{n}samples_bytes"] = total_bytes(rs)
        out[f"unet_forward_{n}samples_kernels"] = len(rs)
        out[f"unet_forward_{n}samples_sum_kernel_ms"] = sum(r.get("us", 0) for r in rs) / 1e3
        # the dominant kernel's dominant shape: the GEGLU projection = the gemm_tc_kernel launches with the longest duration at K = 1280 (60 per forward)
        gem = [r for r in rs if "gemm_tc_kernel<256, 2>" in r["name"]]
        if gem and n == 8:       # the 60 GEGLU projections (M = 8192, N = 10240, K = 1280) are the most populated 10-us duration bucket of this tile shape
            bucket = collections.Counter(round(r["us"] / 10) for r in gem).most_common(1)[0][0]
            top = [r for r in gem if round(r["us"] / 10) == bucket]
            out["geglu_gemm_M8192_bytes"] = sum(r.get("dram__bytes_read", 0) + r.get("dram__bytes_write", 0) for r in top) / len(top)
            out["geglu_gemm_M8192_launches_averaged"] = len(top)
            out["geglu_gemm_M8192_mean_us_under_ncu"] = sum(r["us"] for r in top) / len(top)
    if "decode" in base:
        steps = int(os.environ.get("STEPS", "4"))
        dec = [r for r in rs if any(k in r["name"] for k in ("gemv_mma_kernel", "decode_attn_kernel", "logits_argmax", "store_hidden", "embed_rows"))]
        out["decode_step_bytes"] = total_bytes(dec) / steps
        out["decode_step_sum_kernel_ms"] = sum(r.get("us", 0) for r in dec) / 1e3 / steps
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
