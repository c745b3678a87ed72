# b200-partition-gpu: static one-shot binary in a distroless image (role of reference partition_gpu/Dockerfile:22-27).
FROM gcc:14 AS build
COPY agent/native /src
RUN g++ -O2 -std=c++17 -static -o /b200-partition-gpu /src/partition_gpu.cc
FROM gke.gcr.io/gke-distroless/bash
COPY --from=build /b200-partition-gpu /usr/bin/b200-partition-gpu
CMD ["/usr/bin/b200-partition-gpu", "-logtostderr"]
