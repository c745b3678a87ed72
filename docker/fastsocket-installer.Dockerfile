# Compat image for deploy/transport/compat/fast-socket-installer.yaml: carries Google's fast-socket NCCL net plugin
# (libnccl-net.so) so the DaemonSet's init step can copy it into the host's NCCL directory (role of the reference's
# fast-socket-installer/image/Dockerfile). Inter-node only; on one NVSwitch box libb200coll is the transport.
FROM debian:bookworm-slim
ARG FAST_SOCKET_VERSION=0.0.5
RUN set -eux; \
    apt-get update; \
    apt-get install -y --no-install-recommends ca-certificates curl gnupg; \
    install -d -m 0755 /etc/apt/keyrings; \
    curl -fsSL https://packages.cloud.google.com/apt/doc/apt-key.gpg | gpg --dearmor -o /etc/apt/keyrings/google-cloud.gpg; \
    echo "deb [signed-by=/etc/apt/keyrings/google-cloud.gpg] https://packages.cloud.google.com/apt google-fast-socket main" > /etc/apt/sources.list.d/google-fast-socket.list; \
    apt-get update; \
    apt-get install -y --no-install-recommends "google-fast-socket=${FAST_SOCKET_VERSION}"; \
    rm -rf /var/lib/apt/lists/*
