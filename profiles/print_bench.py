import sys,json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        print(line[:300]); continue
    d=json.loads(line)
    r=d["roofline"]
    print("%s: %.0f GFlop/s  %.2f ms/step frac=%.3f fwd=%s bwd=%s err=%.1e e2e=%s"%(d["config"]["workload"][:12], d["value"], d["ms_per_step"], r["frac"], [round(x,2) for x in r["stage_ms_forward"]], [round(x,2) for x in r["stage_ms_backward"]], d["roundtrip_rel_err"], d.get("e2e",{}).get("value")))
    c=d.get("roofline_hbm_nvlink")
    if c: print("   HBM+NVLink roofline: serial %.2f ms, overlap %.2f ms, measured %.2f ms/transform -> %.2f of serial; NVLink GB/s in exchange stages: %s"%(c["t_roof_serial_ms"], c["t_roof_overlap_ms"], c["t_measured_ms"], c["frac_of_serial_roofline"], c["nvlink_gbs_in_exchange_stages"]))
