for t in ${TARGETS:-0 296 330 370 404 444}; do
  export MFB200_TARGET_CTAS=$t
  python bench.py --steps 32 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('target $t: bench us/launch', round(d['roofline']['us_per_launch'],2), 'us/layer-step', round(d['us_per_layer_step'],2))"
done
