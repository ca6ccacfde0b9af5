import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]); print({k.replace("auxssm_",""):round(v["ms_total"]/v["calls"],3) for k,v in d["kernel_ms"].items()})
    else: print(l.rstrip())
