#!/bin/bash
# Emit a hyper-parameter sweep of 8-GPU training Jobs, one YAML per configuration, under ./jobs/.
# Role of reference demo/gpu-training/generate_job.sh:30-81 (32 ResNet Jobs = 4 depths x 2 batch sizes x 4 learning rates);
# the sweep axes, GPU count and image are parameters here, and the pod requests the b200coll transport profile.
set -eu
OUT="${OUT_DIR:-jobs}"
IMAGE="${IMAGE:-gcr.io/vishnuk-cloud/tf-models-gpu:1.0}"
GPUS="${GPUS_PER_JOB:-8}"
DEPTHS="${DEPTHS:-18 34 50 101}"
BATCHES="${BATCH_SIZES:-64 128}"
LRS="${LEARNING_RATES:-0.01 0.05 0.1 0.5}"
mkdir -p "${OUT}"
n=0
for depth in ${DEPTHS}; do for batch in ${BATCHES}; do for lr in ${LRS}; do
  name="resnet-${depth}-b${batch}-lr$(echo "${lr}" | tr -d .)"
  cat > "${OUT}/${name}.yaml" <<YAML
apiVersion: batch/v1
kind: Job
metadata:
  name: ${name}
  labels: {sweep: resnet}
spec:
  backoffLimit: 1
  template:
    spec:
      restartPolicy: Never
      containers:
      - name: resnet
        image: ${IMAGE}
        command: ["python", "/models/official/resnet/imagenet_main.py"]
        args: ["--resnet_size=${depth}", "--batch_size=${batch}", "--learning_rate=${lr}", "--num_gpus=${GPUS}", "--use_synthetic_data"]
        resources:
          limits:
            nvidia.com/gpu: ${GPUS}
YAML
  n=$((n + 1))
done; done; done
echo "wrote ${n} jobs to ${OUT}/"
