#!/bin/bash
# COS only: enable kdump and reboot once if it was not ready (reference gpudirect-tcpxo/cos-enable-kdump.yaml:59-84, S6).
set -u
HELPER="${KDUMP_HELPER:-/usr/sbin/kdump_helper}"
if ! grep -q "ID=cos" "${OS_RELEASE:-/etc/os-release}"; then echo "not COS, nothing to do"; exit 0; fi
status=$(${HELPER} status 2>&1 || true)
if echo "${status}" | grep -q "kdump is ready"; then echo "kdump already enabled"; exit 0; fi
${HELPER} enable || exit 1
echo "kdump enabled; rebooting to load the crash kernel"
${REBOOT:-systemctl reboot}
