import sys, os, runpy
sys.path.insert(0, os.getcwd())
from isegprobe_b200 import loftup
loftup.LoftUpUpsampler.fuse_ffn = True
sys.argv = ["bench.py", "--workload", "loftup", "--steps", "6", "--warmup", "3", "--no-cpu-baseline", "--no-context"]
runpy.run_path("bench.py", run_name="__main__")
