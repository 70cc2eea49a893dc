for v in libzpack_b200 variant_c variant_b; do
  echo "== $v"
  ZPB_LIB=$PWD/zpack_b200/$v.so python tools/class_bench.py --entries 16384 --groups 8 --classes 1,3,-1 --reps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:200]); continue
    print(d['class'], d['uncomp_GBps'], d['stages'])
"
done
