"""usage: tools/show.py <dir>  -- one line per bench_*.json in a gpurun_out call directory (tools, not product)"""
import glob, json, os, sys
for f in sorted(glob.glob(os.path.join(sys.argv[1], "bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d.get("stages", {}).get("ms") or d.get("stages", {})
        chk = d.get("check") or {}
        print(os.path.basename(f)[6:-5].ljust(22), "%8.2f ms" % d["ms_per_step"],
              {k.replace("dep_", ""): round(v, 2) for k, v in st.items() if isinstance(v, (int, float))} if isinstance(st, dict) else st,
              "ok" if chk.get("ok") else chk, "e2e %.1f ms" % d["e2e"]["ms_per_step"] if d.get("e2e") else "")
    except Exception as e:
        err = f[:-5] + ".err"
        print(os.path.basename(f), "ERR", e, open(err).read()[-300:] if os.path.exists(err) else "")
