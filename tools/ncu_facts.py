"""Record the counters bench.py quotes for the dominant kernel from an `ncu --set full` capture,
together with the SHA-256 of the kernel sources they were measured on:

    python tools/ncu_facts.py gpurun_out/melif_r2i.ncu-rep 444 profiles/r02_melif_ws_final_r2i_ncu_summary.csv

-> profiles/ncu_facts.json["melif_ws_kernel"].  bench.py uses the entry only while the hash of
the sources in the tree equals the recorded one (bench.recorded_ncu_facts); a kernel change
turns `roofline.traffic` and `roofline.issue_bound` into null until a new capture is recorded."""
import csv
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

SOURCES = ["melif.cu", "melif_core.cuh"]


def main(report, notes, profile):
    out = subprocess.run(["ncu", "-i", report, "--csv", "--page", "raw"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units, vals = rows[0], rows[1], rows[-1]

    def metric(name):
        i = head.index(name)
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(units[i], 1.0)
        return float(vals[i].replace(",", "")) * scale
    kernel = vals[head.index("Kernel Name")]
    assert "melif_ws_kernel" in kernel, kernel
    facts = {
        "kernel": kernel[:80], "notes_in_launch": notes, "profile": profile, "sources": SOURCES,
        "source_sha16": bench.source_hash(*SOURCES),
        "dram_bytes_per_note": (metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")) / notes,
        "warp_instructions_per_note": metric("smsp__inst_executed.sum") / notes,
        "shared_wavefronts_per_note": metric("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / notes,
        "duration_us_under_ncu": metric("gpu__time_duration.sum"),
    }
    path = ROOT / "profiles" / "ncu_facts.json"
    data = json.loads(path.read_text()) if path.exists() else {}
    data["melif_ws_kernel"] = facts
    path.write_text(json.dumps(data, indent=1, sort_keys=True) + "\n")
    print(json.dumps(facts, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3])
