import os, sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/tools')
src = open('/root/repo/tools/config_bench.py').read()
src = src.split("# config 3:")[0]
exec(compile(src, 'config_bench_part', 'exec'))
