b() { timeout 200 python bench.py --steps 100 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g  ms %.4f' % (d['value'], d['ms_per_step']))"; }
echo "== base lib"; JR_ENGINE_LIB=josefine_b200/csrc/ab/lib_base.so b
echo "== split1 auto"; JR_ENGINE_LIB=josefine_b200/csrc/ab/lib_split1.so b
echo "== new auto"; b
echo "== new parts=1"; JR_PARTS=1 b
echo "== new parts=4"; JR_PARTS=4 b
echo "== new parts=3"; JR_PARTS=3 b
