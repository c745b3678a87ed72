#!/bin/bash
# Keep systemd-networkd from renaming the host through secondary NICs: set the hostname from metadata and pin a
# DHCP profile that ignores the hostname option on every non-eth0 interface, then reload.
# Behaviour: reference gpudirect-tcpxo/fix-hostname.yaml:29-57.
set -u
MD="${METADATA_URL:-http://metadata.google.internal/computeMetadata/v1}"
NETDIR="${NETWORKD_DIR:-/etc/systemd/network}"
name=$(${CURL:-curl} -sf -H "Metadata-Flavor: Google" "${MD}/instance/hostname" | cut -d. -f1)
if [ -n "${name}" ]; then ${HOSTNAMECTL:-hostnamectl} set-hostname "${name}"; fi
cat > "${NETDIR}/97-temp.network" <<NET
[Match]
Name=!eth0
[Network]
DHCP=yes
[DHCP]
UseHostname=false
NET
${NETWORKCTL:-networkctl} reload
echo "hostname pinned to ${name}"
