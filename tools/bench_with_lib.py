"""Developer tool (GPU): run bench.py against a variant build of the library (tools/micro/build_trace_lab.sh).
    python tools/bench_with_lib.py tools/micro/libusrt_w8.so --steps 50        ("default" = the in-tree library)"""
import sys, os, runpy
sys.path.insert(0, os.getcwd())
from unitysimpleraytracing_b200 import _lib
if sys.argv[1] != "default":
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = ["bench.py"] + sys.argv[2:]
runpy.run_path("bench.py", run_name="__main__")
