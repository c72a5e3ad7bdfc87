"""ICP update() at M = N = 200 000 (bench.py's secondary.icp_200k entry) alone.  usage: python tools/icp200k_once.py"""
import json
import os
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import bench_secondary
from gingr_b200 import api

ctx = api.Context(0)
out = bench_secondary.icp_200k(ctx, root)
for k, v in out["flavours"].items():
    print(k, json.dumps({"update_ms": v["update_ms"], **v["phases_ms"]}))
