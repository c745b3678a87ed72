# b200-nri-device-injector (role of reference nri_device_injector/Dockerfile:22-27).
FROM python:3.12-slim
RUN pip install --no-cache-dir protobuf pyyaml grpcio
COPY container_engine_accelerators_b200 /app/container_engine_accelerators_b200
COPY agent/native/mig_profiles.inc /app/agent/native/mig_profiles.inc
ENV PYTHONPATH=/app
CMD ["python", "-m", "container_engine_accelerators_b200.agent.nri"]
