# Ubuntu / minikube .run driver installer image: ENTRY selects the entrypoint.
FROM ubuntu:24.04
ARG ENTRY=ubuntu
RUN apt-get update && apt-get install -y --no-install-recommends curl ca-certificates kmod build-essential xz-utils bc bison flex libelf-dev libssl-dev && rm -rf /var/lib/apt/lists/*
COPY deploy/driver-installer/lib /opt/driver-installer/lib
COPY deploy/driver-installer/${ENTRY}/entrypoint.sh /opt/driver-installer/${ENTRY}/entrypoint.sh
ENV ENTRY=${ENTRY}
CMD ["/bin/bash", "-c", "/opt/driver-installer/${ENTRY}/entrypoint.sh"]
