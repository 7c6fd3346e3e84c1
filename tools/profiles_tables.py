"""Prints the markdown tables of profiles/README.md (round 2) from the committed JSON files."""
import json
import os

P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def load(name):
    try:
        with open(os.path.join(P, name)) as f:
            return json.load(f)
    except Exception:
        return None


def main():
    print("| GPUs | flat `value` (perm/s) | ms/step | `e2e` (perm/s) | tree leaves | ms/tree | tree perm/s | compute-only ms | exposed all-gather ms | parity |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for n in (1, 2, 4, 8):
        d = load("r2_bench_merkle4_n%d.json" % n)
        if not d:
            continue
        t = d.get("tree", {})
        print("| %d | %.4g | %.3f | %.4g | 4^%s | %.2f | %.4g | %.2f | %.2f | %s |" % (
            n, d["value"], d["ms_per_step"], d["e2e"]["value"], t.get("leaves_log4"), t.get("ms_per_tree", 0), t.get("value", 0),
            t.get("compute_only_ms", 0), t.get("exposed_allgather_ms", 0), t.get("parity")))
    print()
    d = load("r2_bench_merkle4_n8.json")
    if d:
        print("Per-level device times of the 2^28-leaf build, rank 0 (`P252_TIMING` events):\n")
        print("| level | nodes | kernel ms | all-gather ms | gathered MiB | GB/s per rank |")
        print("|---|---|---|---|---|---|")
        for r in d["tree"]["per_level_rank0"]:
            print("| %d | %d | %.3f | %s | %s | %s |" % (r["level"], r["nodes"], r["kernel_ms"], r.get("gather_ms", "–"),
                                                      r.get("gather_MiB", "–"), r.get("gather_GBps", "–")))
        print()
    print("| GPUs | sweep perm/s (all 256 lengths, 2^18 items each per GPU) | s per pass |")
    print("|---|---|---|")
    for n in (1, 2, 4, 8):
        d = load("r2_bench_sweep_n%d.json" % n)
        if d:
            print("| %d | %.4g | %.2f |" % (n, d["value"], d["ms_per_step"] / 1e3))
    d = load("r2_bench_sweep_n1.json") or load("r2_bench_sweep_n8.json")
    if d:
        print("\nPer input length (one GPU's view):\n")
        print("| in_len | perms/item | ms | perm/s |")
        print("|---|---|---|---|")
        tab = {r["in_len"]: r for r in d["sweep"]}
        for L in (1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256):
            if L in tab:
                r = tab[L]
                print("| %d | %d | %.3f | %.4g |" % (L, r["perms_per_item"], r["ms"], r["perm_per_s"]))
        rates = [r["perm_per_s"] for r in d["sweep"]]
        print("\nmin %.4g (in_len %d), max %.4g, mean %.4g over the 256 lengths" % (
            min(rates), d["sweep"][rates.index(min(rates))]["in_len"], max(rates), sum(rates) / len(rates)))
    d = load("r2_small_batch.json")
    if d:
        print("\n| items | lane-split kernel ms | throughput kernel ms | CPU port ms (threads) |")
        print("|---|---|---|---|")
        for r in d["rows"]:
            print("| %d | %.4f | %.4f | %.3f (%d) |" % (r["n"], r["lane_split_ms"], r["throughput_kernel_ms"], r["cpu_port_ms"], r["cpu_threads"]))
    d = load("r2_ncu_summary.json")
    if d:
        print("\n| capture | kernel | grid | duration ms | DRAM read / written MB | regs | warps/SM | fmaheavy % | issue % | fp64 % | top stalls per issue |")
        print("|---|---|---|---|---|---|---|---|---|---|---|")
        for k, v in d.items():
            st = sorted(v.get("stalls_per_issue", {}).items(), key=lambda kv: -kv[1])[:3]
            print("| %s | `%s` | %s | %.3f | %.1f / %.1f | %d | %.1f | %.1f | %.1f | %.1f | %s |" % (
                k, v["kernel"].replace("void ", ""), v["grid"], v["duration_ms"], v["dram_read_MB"], v["dram_write_MB"],
                v["registers_per_thread"], v["warps_active_per_sm"], v["pipe_fmaheavy_active_pct"], v["issue_active_pct"],
                v["inst_pipe_fp64_pct"], ", ".join("%s %.1f" % kv for kv in st)))


if __name__ == "__main__":
    main()
