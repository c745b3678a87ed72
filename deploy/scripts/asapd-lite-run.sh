#!/bin/bash
# asapd-lite supervisor for A4X-Max bare metal: start the daemon with the hairpin probe, and restart the pod when an
# IPv6 duplicate-address-detection failure is seen on a gpu* interface (bounce the link, exit 1).
# Behaviour: reference asapd-lite-installer/asapd-lite-installer-a4x-max-bm-cos.yaml:46-74.
set -u
${ASAPD_RUN:-/run_asapd_lite.sh} --enable-hairpin-probe &
pid=$!
while kill -0 "${pid}" 2>/dev/null; do
  for ifc in $(${IP:-ip} -o link show | awk -F': ' '{print $2}' | grep '^gpu' || true); do
    if ${IP:-ip} -6 addr show dev "${ifc}" | grep -q dadfailed; then
      echo "IPv6 DAD failed on ${ifc}; bouncing the link and restarting"
      ${IP:-ip} link set dev "${ifc}" down; ${IP:-ip} link set dev "${ifc}" up
      kill "${pid}" 2>/dev/null
      exit 1
    fi
  done
  sleep "${ASAPD_POLL_S:-10}"
done
wait "${pid}"
