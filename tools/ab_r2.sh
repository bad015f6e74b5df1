mkdir -p gpurun_out
for rep in 1 2; do
echo "pair+hint";    timeout 300 python tools/step_breakdown.py 60 --only 1 2>&1 | tail -1
echo "pair nohint";  JR_NO_PARTS_HINT=1 timeout 300 python tools/step_breakdown.py 60 --only 1 2>&1 | tail -1
echo "cta+hint";     JR_ENGINE_LIB=$PWD/josefine_b200/csrc/ab/lib_ctabar.so timeout 300 python tools/step_breakdown.py 60 --only 1 2>&1 | tail -1
echo "cta nohint";   JR_NO_PARTS_HINT=1 JR_ENGINE_LIB=$PWD/josefine_b200/csrc/ab/lib_ctabar.so timeout 300 python tools/step_breakdown.py 60 --only 1 2>&1 | tail -1
done
