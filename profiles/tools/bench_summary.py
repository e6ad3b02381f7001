import json, sys
d=json.load(open(sys.argv[1]))
print("ms", d["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["kernel"][:40], d["clocks"]["reasons"])
print("bus", d["bus"]["extra_ms_per_block"], d["bus"]["ms_per_block_mix_plus_bus"])
e=d.get("e2e")
if e:
    print("e2e", e["value"], "pageable", e["pageable_buffers"]["value"], "pinned-in-place", e["pageable_buffers_after_pin_host"]["value"], "link", e["link_peak"]["channel_samples_per_s_if_copies_were_all"], e["fraction_of_link_peak"])
for k,v in (d.get("configs") or {}).items(): print(k, round(v["ms_per_block"],4), v["kernel"], round(v["roofline"]["frac"],3))
if d.get("strong"): print("strong", d["strong"])
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
