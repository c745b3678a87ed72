#!/bin/bash
# Host preparation for GPUDirect-TCPXO, run through `nsenter -at 1`: open the firewall for the data path, load the
# dmabuf import helper, and expose every NIC aperture (PCI 1ae0:0084) under /dev/aperture_devices for LLCM.
# Behaviour: reference gpudirect-tcpxo/nccl-tcpxo-installer.yaml:49-70.
set -u
SYS_PCI="${SYS_PCI_DEVICES:-/sys/bus/pci/devices}"
APERTURE_DIR="${APERTURE_DIR:-/dev/aperture_devices}"
${IPTABLES:-/sbin/iptables} -I INPUT -p tcp -m tcp -j ACCEPT
${MODPROBE:-/sbin/modprobe} import-helper
mkdir -p "${APERTURE_DIR}"
for dev in "${SYS_PCI}"/*; do
  [ -r "${dev}/vendor" ] && [ -r "${dev}/device" ] || continue
  if [ "$(cat "${dev}/vendor")" = "0x1ae0" ] && [ "$(cat "${dev}/device")" = "0x0084" ]; then
    bdf=$(basename "${dev}")
    mkdir -p "${APERTURE_DIR}/${bdf}"
    ${MOUNT:-mount} --bind "${dev}" "${APERTURE_DIR}/${bdf}"
    chmod 666 "${APERTURE_DIR}/${bdf}"/resource* 2>/dev/null || true
    echo "exposed aperture ${bdf}"
  fi
done
