import json,sys
d=json.load(open(sys.argv[1]))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
for k,v in d["per_op"].items(): print(k, {a:(round(b,3) if b else b) for a,b in v.items()})
for k,v in d["kernels"].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
print({k:v for k,v in d["roofline"].items() if k!="traffic_source"})
