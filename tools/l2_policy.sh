for pol in 2 1 0; do
S3D_L2_POLICY=$pol python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/bp$pol.json
python -c "
import json; d=json.load(open('gpurun_out/bp$pol.json')); print('policy',$pol, round(d['value']), {k:round(v['ms']/max(v['launches'],1)*1e3,1) for k,v in d['roofline']['kernels'].items()}, d['config']['chunk_retries'])"
done
