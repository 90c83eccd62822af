for B in 16 256; do
for v in "EGN_TC_V3=0" "EGN_TC_V3=0 EGN_TC_V4_SW=128"; do
echo "== batch $B $v"
env $v EGN_TC_TS=1 EGN_TC_TS_DUMP=1 EGN_TC_VERBOSE=1 python tools/layer_bench.py --child --batch $B --iters 1 --dtype 2 --first 1 --shapes 3 2>&1 | grep -E "egn-ts4|RESULT|v4-tapwin" | cut -c1-260
done; done
