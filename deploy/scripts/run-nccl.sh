#!/bin/bash
# run-nccl.sh <benchmark> <ld_library_path> <gpus_per_node> <data_nics> <min> <max> <nhosts>
# One rank per GPU (-g 1), nccl-tests protocol `-b min -e max -f 2 -w 5 --iters 100 -c 0`, TCPX environment.
# Behaviour: reference gpudirect-tcpx/nccl-config.yaml:18-63.
set -eu
BENCH=$1; LDP=$2; GPN=$3; NICS=$4; MINB=$5; MAXB=$6; NHOSTS=$7
NP=$(( GPN * NHOSTS ))
mpirun --mca btl tcp,self --mca btl_tcp_if_include eth0 --allow-run-as-root -np "${NP}" --hostfile "/scripts/hostfiles${NHOSTS}/hostfile${GPN}" \
  -x LD_LIBRARY_PATH="${LDP}" -x NCCL_SOCKET_IFNAME=eth0 -x NCCL_ALGO=Ring -x NCCL_PROTO=Simple -x NCCL_CROSS_NIC=0 \
  -x NCCL_NET_GDR_LEVEL=PIX -x NCCL_P2P_PXN_LEVEL=0 -x NCCL_MAX_NCHANNELS=8 -x NCCL_MIN_NCHANNELS=8 -x NCCL_BUFFSIZE=4194304 \
  -x NCCL_P2P_NVL_CHUNKSIZE=1048576 -x NCCL_P2P_PCI_CHUNKSIZE=524288 -x NCCL_P2P_NET_CHUNKSIZE=524288 -x NCCL_DYNAMIC_CHUNK_SIZE=524288 \
  -x NCCL_GPUDIRECTTCPX_SOCKET_IFNAME="${NICS}" -x NCCL_GPUDIRECTTCPX_CTRL_DEV=eth0 -x NCCL_GPUDIRECTTCPX_FORCE_ACK=0 \
  -x NCCL_GPUDIRECTTCPX_TX_COMPLETION_NANOSLEEP=100 -x NCCL_GPUDIRECTTCPX_PROGRAM_FLOW_STEERING_WAIT_MICROS=1000000 \
  -x NCCL_NSOCKS_PERTHREAD=4 -x NCCL_SOCKET_NTHREADS=1 -x NCCL_DEBUG=INFO -x NCCL_DEBUG_SUBSYS=ENV -x CUDA_VISIBLE_DEVICES=0,1,2,3,4,5,6,7 \
  taskset -c 32-63 "/third_party/nccl-tests-mpi/build/${BENCH}" -b "${MINB}" -e "${MAXB}" -f 2 -g 1 -w 5 --iters 100 -c 0 2>&1 | tee "/tmp/${BENCH}_${NP}.txt"
