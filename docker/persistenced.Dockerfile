# b200-persistenced sidecar (role of reference nvidia-persistenced-installer/Dockerfile:28-38: needs ldconfig in the image).
FROM gcc:14 AS build
COPY agent/native /src
RUN g++ -O2 -std=c++17 -static -o /b200-persistenced /src/persistenced.cc
FROM debian:12-slim
COPY --from=build /b200-persistenced /usr/bin/b200-persistenced
CMD ["/usr/bin/b200-persistenced", "-logtostderr"]
