import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"ERR",e); continue
    print(f,"N",d["n_gpus"],"value %.3f"%d["value"],"e2e %.3f"%d["e2e"]["value"],"parity",d["parity"],"sha",d["proof_sha"],"frac %.3f"%d["roofline"]["frac"])
    print("   ",{k:round(v,2) for k,v in d["breakdown_ms"].items()}, d.get("host_phases_ms"))
