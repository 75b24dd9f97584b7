set -e
L=rsoccer_isaac_cleanrl_b200
cp $L/libvss_b200.so /tmp/A.so; cp $L/libvss_b200_B.so /tmp/B.so
for v in A B A B; do
  cp /tmp/$v.so $L/libvss_b200.so
  timeout 150 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-e2e --no-sweep --skip-reference-torch --ppo-updates 12 --ppo-legs sa:4096 > /tmp/o.json 2>/dev/null
  python - <<EOF
import json
d=json.loads(open("/tmp/o.json").read().strip().splitlines()[-1])
l=d["ppo_legs"][0]; print("$v", "steady %.4e"%l["sps_steady"], "rollout %.3f ms"%(l["rollout_s_steady"]*1e3), "update %.3f ms"%(l["update_s_steady"]*1e3))
EOF
done
cp /tmp/A.so $L/libvss_b200.so
