import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    r = d["roofline"]
    rnd = lambda x, n: None if x is None else round(x, n)
    print(f.split("/")[-1], "it/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "heavy", rnd(r["avg_launch_ms"], 4),
          "frac", rnd(r["frac"], 3), "iterfrac", round(r["iteration"]["frac"], 3), "FT", d["plan"]["tile_freqs"], "items", d["plan"]["nitems"],
          "clk", d["clocks"]["sm_mhz"], "e2e", round(d["e2e"]["value"], 1), "loss", d["loss_first_last"])
