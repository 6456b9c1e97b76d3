for st in ${STAGES:-0 1 2 3 4 5}; do
S3D_DEBUG_STAGE=$st python bench.py --no-cpu-baseline --no-e2e --steps 4 2>&1 | tail -1 > gpurun_out/b$st.json
python -c "
import json; d=json.load(open('gpurun_out/b$st.json')); print($st, round(d['value']), {k:round(v['ms']/v['launches']*1e3,1) for k,v in d['roofline']['kernels'].items()})"
done
