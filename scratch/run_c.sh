cp dnpy_b200/libdnpy_b200.so /tmp/lib_tr8.so
cp scratch/libdnpy_tr4.so /tmp/lib_tr4.so; cp scratch/libdnpy_tr6.so /tmp/lib_tr6.so
for tr in 8 4 6; do
cp /tmp/lib_tr$tr.so dnpy_b200/libdnpy_b200.so
echo "=== TR $tr"
SVMATH=bf16 python scratch/sv_bench.py 0 2>&1 | grep debug
timeout 300 python bench.py --no-cpu --math bf16 > /tmp/b.json 2>/tmp/b.err
python - <<PY
import json
d=json.load(open('/tmp/b.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'])
for t in d['top_kernels']:
    if '_sv' in t['kernel']: print(f"{t['ms_per_step']*1e3:8.1f} us {t['kernel']}")
PY
done
