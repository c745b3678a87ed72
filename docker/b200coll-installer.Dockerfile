# b200coll-installer: the transport payload image (role of reference fast-socket-installer/image/Dockerfile:1-7 and the opaque
# nccl-plugin-* images): libb200coll.so, the NCCL-API shim, the perf tool, the env profile and the tuner table under /opt/b200coll.
FROM nvidia/cuda:12.9.1-devel-ubuntu22.04 AS build
COPY coll /src/coll
COPY tools /src/tools
RUN make -C /src/coll && make -C /src/tools
FROM nvidia/cuda:12.9.1-base-ubuntu22.04
COPY --from=build /src/coll/lib/libb200coll.so /src/coll/lib/libb200coll_nccl.so /opt/b200coll/lib/
COPY --from=build /src/build/b200coll_perf /src/build/mps_probe /opt/b200coll/bin/
COPY coll/tuner/b200_nvswitch.tbl /opt/b200coll/tuner/b200_nvswitch.tbl
COPY deploy/scripts/b200coll-env-profile.sh /opt/b200coll/b200coll-env-profile.sh
COPY deploy/scripts/b200coll-install.sh /scripts/install.sh
CMD ["/scripts/install.sh"]
