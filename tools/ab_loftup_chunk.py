"""A/B of LoftUpUpsampler.chunk_images on the headline workload: python tools/ab_loftup_chunk.py 8"""
import sys, os, runpy
sys.path.insert(0, os.getcwd())
from isegprobe_b200 import loftup
loftup.LoftUpUpsampler.chunk_images = int(sys.argv[1])
sys.argv = ["bench.py", "--workload", "loftup", "--steps", "6", "--warmup", "3", "--no-cpu-baseline", "--no-context"]
runpy.run_path("bench.py", run_name="__main__")
