import sys, os, runpy
sys.path.insert(0, os.getcwd())
from unitysimpleraytracing_b200 import _lib
if sys.argv[1] != "default": _lib.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = ["bench.py"] + sys.argv[2:]
runpy.run_path("bench.py", run_name="__main__")
