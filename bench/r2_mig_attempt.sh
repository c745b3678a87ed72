#!/bin/bash
# BASELINE config 5 on real hardware, as far as the box allows: switch GPU 0 to MIG mode, let b200-partition-gpu cut it into seven
# 1g.23gb slices, have the device plugin advertise them, then put everything back. Every step is logged; whatever happens, the EXIT trap
# destroys the instances and disables MIG mode again, and the final state is recorded. The reboot the partitioner would request after
# enabling MIG mode is redirected to a hook file (never signal pid 1 on a shared box).
O=gpurun_out/r2_mig; mkdir -p gpurun_out
SMI=$(command -v nvidia-smi)
state() { $SMI -i 0 --query-gpu=mig.mode.current,mig.mode.pending --format=csv,noheader 2>&1; }
cleanup() {
  { echo "--- cleanup"; $SMI mig -i 0 -dci; $SMI mig -i 0 -dgi; timeout 60 $SMI -i 0 -mig 0; echo "final: $(state)"; } >> ${O}_log.txt 2>&1
  echo "final MIG state: $(state)"
}
trap cleanup EXIT
echo "before: $(state)" | tee ${O}_log.txt
timeout 90 $SMI -i 0 -mig 1 >> ${O}_log.txt 2>&1; echo "nvidia-smi -mig 1 rc=$?" | tee -a ${O}_log.txt
echo "after enable: $(state)" | tee -a ${O}_log.txt
case "$(state)" in
  Enabled,\ Enabled*) ;;
  *) echo "RESULT: MIG mode could not be enabled in place on this box (state '$(state)'): refusal recorded, nothing partitioned" | tee -a ${O}_log.txt; exit 0;;
esac
cat > /tmp/gpu_config.json <<'J'
{"GPUPartitionSize": "1g.23gb"}
J
B200_PARTITION_REBOOT_HOOK=/tmp/reboot.hook timeout 180 ./build/agent/b200-partition-gpu -gpu-config /tmp/gpu_config.json -nvidia-smi-path $SMI >> ${O}_log.txt 2>&1; echo "b200-partition-gpu rc=$? reboot-hook=$(cat /tmp/reboot.hook 2>/dev/null)" | tee -a ${O}_log.txt
$SMI mig -lgi >> ${O}_log.txt 2>&1
$SMI -L | tee -a ${O}_log.txt
echo "GPU instances: $($SMI mig -lgi | grep -c 'MIG 1g')" | tee -a ${O}_log.txt
ls -l /proc/driver/nvidia/capabilities/gpu0/mig 2>&1 | head -20 >> ${O}_log.txt
# the device plugin in MIG mode on the real /dev, /proc, NVML: must advertise seven nvidia0/gi<N>
python - >> ${O}_log.txt 2>&1 <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
from conformance import run as conf
n = conf.Node(conf.PRESETS["native"], real=True, config={"GPUPartitionSize": "1g.23gb"})
try:
    c = n.connect(30)
    stream, devs = conf.first_list(c)
    ids = sorted(devs)
    print("plugin advertises:", ids)
    gi = [i for i in ids if i.startswith("nvidia0/gi")]
    print("RESULT: device plugin lists", len(gi), "MIG resources on GPU 0")
    if gi:
        cr = c.allocate([gi[0]]).container_responses[0]
        print("allocate", gi[0], "->", [d.host_path for d in cr.devices])
    stream.cancel()
except Exception as e:
    print("plugin in MIG mode failed:", repr(e)); print(n.logs()[-3000:])
finally:
    n.close()
PY
grep RESULT ${O}_log.txt
