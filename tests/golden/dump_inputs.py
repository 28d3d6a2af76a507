"""Writes the input of every known-answer row (tests/golden/encode_rows.json) to <dir>/<kind>-<size>-<seed>.bin
so that the C# xUnit project bindings/csharp/ParityDump can replay them against the real K4os engine.
    python tests/golden/dump_inputs.py /tmp/k4_inputs
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import inputs  # noqa: E402


def main(out):
    os.makedirs(out, exist_ok=True)
    rows = json.load(open(os.path.join(HERE, "encode_rows.json")))["rows"]
    for r in rows:
        open(os.path.join(out, f"{r['kind']}-{r['size']}-{r['seed']}.bin"), "wb").write(
            inputs.gen(r["kind"], r["size"], r["seed"]))
    print(len(rows), "inputs written to", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "k4_inputs")
