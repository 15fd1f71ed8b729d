#!/usr/bin/env python
"""Turn the JSON lines of tools/tune_emitters.py into the markdown table of profiles/r02_emitters.md.

    python tools/summarise_tune.py gpurun_out/<tag>_tune.jsonl > table.md"""
import json
import sys

SHAPES = {0: "4,2,2", 1: "2,2,2", 2: "4,2,4", 3: "4,4,2 (3 CTAs/SM)", 4: "4,4,4 (3 CTAs/SM)", 5: "8,2,4 (3 CTAs/SM)"}


def main(path):
    rows = [json.loads(line) for line in open(path)]
    print("| workload (envs) | path | emitter | shape | split | rings | us/step | TB/s | frac |")
    print("|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        if "error" in r:
            print(f"| {r.get('workload')} | | {r.get('emit')} | {r.get('shape')} | | | error: {r['error'][:60]} | | |")
        elif "emit" not in r:
            print(f"| {r['workload']} ({r['batch']}) | {r['path']} | | | | | {r['us_per_step']} | | |")
        else:
            shape = SHAPES.get(r["shape"], r["shape"]) if r["emit"] == "image" else "-"
            print(f"| {r['workload']} ({r['batch']}) | {r['path']} | {r['emit']} | {shape} | {r.get('specialised', '-')} | {r.get('ring', '-')} | "
                  f"{r['us_per_step']} | {r['tbs']} | {r['frac']} |")


if __name__ == "__main__":
    main(sys.argv[1])
