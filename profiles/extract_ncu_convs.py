"""Per-launch summary of an `ncu --set full` capture of k_conv_mma launches (profiles/ncu_full.sh).
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python profiles/extract_ncu_convs.py raw.csv out.json [traffic.json]
Layers are recognised by (block size, shared memory, duration); the gate convolutions are the launches over 150 us."""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return None
out = []
for r in rows[2:]:
    m = re.search(r"k_\w+(<\d+>)?", r[ix["Kernel Name"]])
    e = {"kernel": m.group(0) if m else r[ix["Kernel Name"]],
         "block": int(f(r, "launch__block_size")), "regs": int(f(r, "launch__registers_per_thread")),
         "smem_dyn_KB": round(f(r, "launch__shared_mem_per_block_dynamic"), 1),
         "time_us": round(f(r, "gpu__time_duration.sum"), 1) if hdr and rows[1][ix["gpu__time_duration.sum"]] in ("us", "usecond") else f(r, "gpu__time_duration.sum"),
         "tensor_pct": round(f(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active") or 0, 1)
                       if "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active" in ix else None,
         "dram_read_MB": round((f(r, "dram__bytes_read.sum") or 0) * {"Gbyte": 1e3, "Mbyte": 1, "Kbyte": 1e-3, "byte": 1e-6}.get(rows[1][ix["dram__bytes_read.sum"]], 1), 1),
         "dram_write_MB": round((f(r, "dram__bytes_write.sum") or 0) * {"Gbyte": 1e3, "Mbyte": 1, "Kbyte": 1e-3, "byte": 1e-6}.get(rows[1][ix["dram__bytes_write.sum"]], 1), 1),
         "lts_pct": round(f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed") or 0, 1),
         "l1tex_pct": round(f(r, "l1tex__throughput.avg.pct_of_peak_sustained_active") or 0, 1),
         "sm_pct": round(f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed") or 0, 1),
         "inst_M": round((f(r, "smsp__inst_executed.sum") or 0) / 1e6, 2)}
    out.append(e)
ORDER = ["scratch0", "scratch1", "masks0", "masks1", "enc0", "lstm0", "enc1", "lstm1", "enc2", "lstm2", "dec0", "lstm3", "dec1", "lstm4", "dec2"]
ORDER2 = ["enc0", "lstm0", "enc1", "lstm1", "enc2", "lstm2", "dec0", "lstm3", "dec1", "lstm4", "dec2", "heads0", "scratch1", "masks1"]
if "--step-order-r02" in sys.argv:    # round 2: 14 convolutions per cell step, capture starts at enc0 (profiles/r02_call2.sh, skip 624)
    for i, e in enumerate(out): e["layer"] = ORDER2[i % len(ORDER2)]
if "--step-order" in sys.argv:        # the capture starts at the first head conv of a cell step (profiles/ncu_full.sh, skip 640)
    for i, e in enumerate(out): e["layer"] = ORDER[i % len(ORDER)]
json.dump({"capture": sys.argv[1], "launches": out}, open(sys.argv[2], "w"), indent=1)
gate = [e for e in out if e["time_us"] and e["time_us"] > 130 and (e["kernel"].endswith("<2>") or "k_conv_pair" in e["kernel"])]   # wide-path / pair launches: the conv-LSTM gate convs
for e in out: print(e)
if len(sys.argv) > 3 and gate:
    mean = sum((e["dram_read_MB"] + e["dram_write_MB"]) for e in gate) / len(gate) * 1e6
    json.dump({"kernel": "k_conv_mma gate convolutions (5 launches per cell step)", "dram_bytes_per_launch": mean,
               "source": "%s (dram__bytes_read.sum + dram__bytes_write.sum, mean over the %d gate-conv launches of the capture)" % (sys.argv[2], len(gate))},
              open(sys.argv[3], "w"), indent=1)
    print("gate convs:", len(gate), "mean DRAM bytes/launch %.1f MB" % (mean / 1e6))
