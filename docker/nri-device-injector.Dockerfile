# b200-nri-device-injector: static C++ binary on distroless (role of reference nri_device_injector/Dockerfile:22-27, a CGO-off Go binary).
FROM gcc:14 AS build
COPY agent/native /src
RUN g++ -O2 -std=c++17 -static /src/dp/nri_injector.cc -o /b200-nri-device-injector -lpthread
FROM gke.gcr.io/gke-distroless/bash
COPY --from=build /b200-nri-device-injector /usr/bin/b200-nri-device-injector
CMD ["/usr/bin/b200-nri-device-injector"]
