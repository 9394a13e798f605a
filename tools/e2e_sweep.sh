for ks in 1 2 3 4; do for c in 16 17 18; do
MIRO_GPU_KSTREAMS=$ks MIRO_GPU_CHUNK=$c python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('kstreams $ks chunk $c e2e %7.1f' % d['e2e']['value'])"
done; done
MIRO_GPU_KSTREAMS=3 MIRO_GPU_CHUNK=17 python tools/e2e_timeline.py 2>&1 | tail -20
