"""A/B helper: run bench.py's timed loop in-process under module-level switches.  Usage: python tools/ab_step.py name=value ...
(names: HIGH_PRIORITY_MAIN, WGRAD_OVERLAP, SIDE_STREAM_PRIORITY); prints ms per step of the resident-input loop."""
import json
import io
import os
import runpy
import sys
import contextlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semisupervisedobjectdetection_b200 import ops
from semisupervisedobjectdetection_b200 import teacher_student as TS

for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "HIGH_PRIORITY_MAIN":
        TS.HIGH_PRIORITY_MAIN = bool(int(v))
    elif k == "WGRAD_OVERLAP":
        ops.WGRAD_OVERLAP = bool(int(v))
    elif k == "SIDE_STREAM_PRIORITY":
        ops.SIDE_STREAM_PRIORITY = int(v)
args = sys.argv[1:]
sys.argv = ["bench.py", "--steps", "10", "--warmup", "3", "--no-cpu-baseline"]
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    try:
        runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"), run_name="__main__")
    except SystemExit:
        pass
for line in buf.getvalue().splitlines():
    if line.startswith("{"):
        d = json.loads(line)
        print(" ".join(args), "->", round(d["ms_per_step"], 2), "ms/step | e2e", round(d["e2e"]["ms_per_step"], 2), "| serial", round(d["roofline"]["serial_step_ms"], 2))
