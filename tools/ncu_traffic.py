"""DRAM traffic per launch of the dominant kernel family (conv_halo*_kernel) from an ncu metrics list of bench.py's own command:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_halo \
        -c 190 --csv --log-file gpurun_out/r2_halo_dram.csv python bench.py --steps 2 --warmup 3 --no-sustained --no-cfg3 --no-cpu
    python tools/ncu_traffic.py gpurun_out/r2_halo_dram.csv 38 "<command>"     ->  profiles/r2_traffic.json (read by bench.py)

38 = halo launches of one cfg-2 step; the average is taken over whole steps only.  A 4th argument "drop-last" also reports the last
halo launch of every step on its own: out.conv, whose epilogue carries the fused posterior update (x_t read, x_{t-1} fp32 +
16-bit written): "fused_out_conv_bytes_per_launch", and the family average without it."""
import collections, csv, json, os, re, sys

src, per_step, cmd = sys.argv[1], int(sys.argv[2]), sys.argv[3]
drop_last = len(sys.argv) > 4 and sys.argv[4] == "drop-last"
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
h = rows[0]
ii, ki, mi, ui, vi = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
L = collections.OrderedDict()
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
for r in rows[1:]:
    d = L.setdefault(r[ii], {"kernel": re.sub(r"\(.*$", "", r[ki])})
    d[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
launches = list(L.values())
n = len(launches) // per_step * per_step
sel = launches[len(launches) - n:]   # the last whole steps (the first launches of a process include cold weights / JIT effects)
fused = [d for i, d in enumerate(sel) if i % per_step == per_step - 1] if drop_last else []
plain = [d for i, d in enumerate(sel) if i % per_step != per_step - 1] if drop_last else sel
tot = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in sel)
us = sum(d["gpu__time_duration.sum"] for d in sel)
out = {"halo_bytes_per_launch": tot / n, "halo_launches_averaged": n, "halo_avg_us_under_ncu": us / n,
       "halo_read_bytes_per_launch": sum(d["dram__bytes_read.sum"] for d in sel) / n,
       "halo_write_bytes_per_launch": sum(d["dram__bytes_write.sum"] for d in sel) / n,
       **({"halo_bytes_per_launch_without_fused_out_conv": sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in plain) / len(plain),
           "fused_out_conv_bytes_per_launch": sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in fused) / len(fused),
           "fused_out_conv_avg_us_under_ncu": sum(d["gpu__time_duration.sum"] for d in fused) / len(fused)} if fused else {}),
       "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over {n} conv_halo*_kernel launches ({n // per_step} cfg-2 steps) of: {cmd}"}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, "profiles", "r2_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
