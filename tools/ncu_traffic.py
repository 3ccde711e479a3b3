"""profiles/<round>_dram_traffic_per_launch.json from a summarised ncu --set full capture (tools/summarize_ncu.py CSV):
dram__bytes_read.sum + dram__bytes_write.sum averaged over each kernel's launches. bench.py reads it for roofline.traffic.
  python tools/ncu_traffic.py profiles/r2_ncu_full_vgg_batch8.csv profiles/r2_dram_traffic_per_launch.json"""
import csv
import json
import re
import sys

src, out = sys.argv[1], sys.argv[2]
acc = {}
for r in csv.DictReader(open(src)):
    name = re.sub(r"<.*", "", r["kernel"].replace("void ", "").replace("unnamed>::", "").strip())
    try:
        b = (float(r["dram_rd_MB"]) + float(r["dram_wr_MB"])) * 1e6
    except ValueError:
        continue
    a = acc.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += b
    try:
        a[2] += float(r["tensor_pct"])
    except ValueError:
        pass
json.dump({"source": "%s (ncu --set full --clock-control none, one cold pass of the 13 VGG16 layers x 3 ops at batch 8): dram__bytes_read.sum + "
                     "dram__bytes_write.sum averaged over each kernel's launches" % src,
           "kernels": {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / v[0], "tensor_pipe_pct_avg": v[2] / v[0]} for k, v in acc.items()}},
          open(out, "w"), indent=1)
print(open(out).read()[:1500])
