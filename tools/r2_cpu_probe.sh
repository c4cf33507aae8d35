cd /root/repo
python - <<'PY'
import sys, os
sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import swco, helpers as H
from multiprocessing import Pool
def mk(i): return H.raw_deflate(H.textlike(65536, 2+i))
with Pool(32) as p: units = p.map(mk, range(256))
lz = [H.lz4_block_compress(H.textlike(65536, 2+i)) for i in range(64)]
for th in (1,2,4,8,16,32,64,128):
    sec,nb,f = swco.batch_mt("deflate", units, 200*th, th)
    s2,n2,f2 = swco.batch_mt("lz4_block", lz, 4000*th, th)
    print(th, 'deflate GB/s', round(nb/sec/1e9,3), 'per-thread', round(nb/sec/1e9/th,4), '| lz4 GB/s', round(n2/s2/1e9,2), flush=True)
print(open('/proc/cpuinfo').read().count('processor\t'), os.cpu_count())
os.system("lscpu | grep -E 'Model name|Socket|Core|Thread|NUMA|MHz' ")
PY
