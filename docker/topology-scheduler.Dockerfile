# Topology scheduler + labeler (the reference ships these as ConfigMap-mounted scripts on python:3.10, gke-topology-scheduler/README.md:30-32).
FROM python:3.12-slim
RUN pip install --no-cache-dir requests pyyaml
COPY container_engine_accelerators_b200 /app/container_engine_accelerators_b200
COPY agent/native/mig_profiles.inc /app/agent/native/mig_profiles.inc
ENV PYTHONPATH=/app
CMD ["python", "-m", "container_engine_accelerators_b200.scheduler.daemon"]
