# b200-device-plugin (native): one static C++ binary in a distroless image — the footprint of the reference's Go plugin
# (reference Dockerfile:15-36 builds a cgo binary onto distroless/base; envelope 50m CPU / 100Mi, cmd/nvidia_gpu/device-plugin.yaml:47-52).
FROM gcc:14 AS build
COPY agent/native /src
RUN g++ -O2 -std=c++17 -static-libstdc++ -static-libgcc /src/dp/device_plugin.cc /src/b200agent_nvml.cc -o /b200-device-plugin -ldl -lpthread
FROM gcr.io/distroless/base
COPY --from=build /b200-device-plugin /usr/bin/b200-device-plugin
# distroless/base ships libssl/libcrypto, which the Kubernetes API client dlopens for https (agent/native/dp/kube.hpp).
CMD ["/usr/bin/b200-device-plugin", "-logtostderr", "-enable-container-gpu-metrics", "-enable-health-monitoring", "-publish-driver-version"]
