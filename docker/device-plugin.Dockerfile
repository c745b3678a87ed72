# b200-device-plugin: Python agent + native NVML binding (role of reference Dockerfile:15-36).
FROM python:3.12-slim AS build
RUN apt-get update && apt-get install -y --no-install-recommends g++ make && rm -rf /var/lib/apt/lists/*
COPY agent/native /src/agent/native
RUN make -C /src/agent/native ../../build/agent/libb200agent_nvml.so
FROM python:3.12-slim
RUN pip install --no-cache-dir grpcio protobuf prometheus_client pyyaml requests
COPY container_engine_accelerators_b200 /app/container_engine_accelerators_b200
COPY agent/native/mig_profiles.inc /app/agent/native/mig_profiles.inc
COPY --from=build /src/build/agent/libb200agent_nvml.so /usr/local/lib/libb200agent_nvml.so
ENV PYTHONPATH=/app B200AGENT_NATIVE_LIB=/usr/local/lib/libb200agent_nvml.so
CMD ["python", "-m", "container_engine_accelerators_b200.agent.main", "--enable-container-gpu-metrics", "--enable-health-monitoring"]
