"""Summarise an .ncu-rep (read offline with `ncu -i`): selected raw metrics, the warp-stall
breakdown of the source page aggregated by reason, and the executed-instruction mix.
usage: python tools/ncu_summary.py <file.ncu-rep> <out.json> [raw.csv]"""
import collections, csv, json, re, subprocess, sys

KEEP = re.compile(
    r"gpu__time_duration|dram__bytes|dram_throughput|lts__t_sector_hit|lts__throughput|l1tex__data_pipe_lsu_wavefronts"
    r"|l1tex__data_bank_conflicts_pipe_lsu_mem_shared|sm__inst_executed_pipe_(fp64|alu|fma|lsu|xu|tensor|uniform)"
    r"|sm__pipe_(fp64|tensor|alu|fma)\w*cycles_active|smsp__issue_active|smsp__inst_executed\.sum|sm__warps_active"
    r"|launch__(registers|shared|block|grid|occupancy)|smsp__average_warps_issue_stalled|sm__throughput|sm__cycles_elapsed"
    r"|smsp__sass_thread_inst_executed_op_d(add|fma|mul)_pred_on\.sum|local_op_(ld|st)\.sum|imma|tmem|sm__clock")


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    res = {"source": rep, "kernel": vals[hdr.index("Kernel Name")], "block": vals[hdr.index("Block Size")],
           "grid": vals[hdr.index("Grid Size")], "metrics": {}}
    for h, u, v in zip(hdr, units, vals):
        if KEEP.search(h) and v not in ("", "n/a"):
            res["metrics"][h] = (v + " " + u).strip()
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w", newline="") as f:
            csv.writer(f).writerows(raw)
    src = page(rep, "source")
    hi = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if hi:
        h = src[hi[0]]
        ix = {k: i for i, k in enumerate(h)}
        stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        tot, byop, mix = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
        for r in src[hi[0] + 1:]:
            if len(r) < len(h):
                continue
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
            op = m.group(2) if m else "?"
            op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDTM", "STTM", "SHFL", "IMAD", "UTC")) else op.split(".")[0]
            mix[op] += int(r[ix["Instructions Executed"]] or 0)
            for s in stalls:
                v = int(r[ix[s]] or 0)
                tot[s] += v
                byop[s][op] += v
        total = sum(tot.values()) or 1
        res["warp_stall_samples_pct"] = {s: round(100.0 * v / total, 2) for s, v in tot.most_common() if v}
        res["warp_stall_top_opcodes"] = {s: dict(byop[s].most_common(5)) for s, _ in tot.most_common(8)}
        n = sum(mix.values()) or 1
        res["warp_instructions_executed"] = n
        res["instruction_mix_pct"] = {k: round(100.0 * v / n, 2) for k, v in mix.most_common(24)}
    json.dump(res, open(dst, "w"), indent=1)
    print(dst, res["kernel"][:60], res["metrics"].get("gpu__time_duration.sum"))


if __name__ == "__main__":
    main()
