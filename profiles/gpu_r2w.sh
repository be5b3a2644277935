for S in 1 2; do
XPCS_FIN_SEGS=$S timeout 300 python bench.py --no-cpu --no-e2e --steps 5 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); k=j['kernels']; print('segs $S ms/step %.3f'%j['ms_per_step'], {x:round(k[x]['ms_per_step'],3) for x in ('k_finalize','k_scatter_rec','k_hist','k_slice_scan','k_slice_len')})"
done
