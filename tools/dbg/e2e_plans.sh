for plan in "1,1" "1,2,3,3,3.5,3.5" "1,2,4,8" "1,2,4,4,4" "1,3,6,6" "1,1,2,2,3,3,4" "1,2,2,3,3,3,3" ; do
SWB200_TRACE=1 SWB200_CHUNK_PLAN=$plan python bench.py --steps 4 --warmup 3 2>gpurun_out/tr.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(\"plan $plan e2e\", round(j[\"e2e\"][\"value\"]), round(j[\"e2e\"][\"ms_per_step\"],2))"
grep "TRACE plan" gpurun_out/tr.err | tail -1
grep "TRACE lane" gpurun_out/tr.err | tail -2 | sed -E 's/ counters@[0-9.]+//g' | cut -c1-400
done
