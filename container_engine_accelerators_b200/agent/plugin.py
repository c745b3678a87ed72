"""kubelet DevicePlugin v1beta1 service + Registration client, over real gRPC on Unix sockets.

Contract: reference pkg/gpu/nvidia/beta_plugin.go:35-145:
  * GetDevicePluginOptions -> empty options (so the kubelet never calls PreStart / GetPreferredAllocation)
    unless --preferred-allocation-policy is set: then get_preferred_allocation_available = true and
    GetPreferredAllocation answers with agent/preferred.py (NUMA-aligned, sharing-aware; new in this build)
  * ListAndWatch -> full list, then the full list again on every health change
  * Allocate -> per container: requested device specs + default devices (nvidiactl, nvidia-uvm, [modeset,
    uvm-tools]) all "mrw" with HostPath == ContainerPath; mounts; MPS envs; whole RPC fails on the first bad id
  * Register{Version: v1beta1, Endpoint: <socket basename>, ResourceName}, no Options
New: when GPUConfig.Transport == "b200coll", Allocate also exports the collective library's env profile
(the lib dir is already covered by the /usr/local/nvidia mount) — SURVEY §5.8-7.
Fix over the reference: ListAndWatch observes stream cancellation and server stop (beta_plugin.go:44-53 never returns).
"""
from __future__ import annotations

import logging
import queue

import grpc

from . import preferred, protos, sharing, transport
from .protos import deviceplugin as pb

log = logging.getLogger("b200-device-plugin")


def _device_msg(dev) -> "pb.Device":
    d = pb.Device(ID=dev.id, health=dev.health)
    if dev.numa_node is not None:
        d.topology.nodes.add(ID=int(dev.numa_node))
    return d


class DevicePluginService:
    def __init__(self, manager):
        self.ngm = manager

    # ------------------------------------------------------------------ handlers
    @property
    def policy(self) -> str:
        return getattr(self.ngm, "preferred_allocation_policy", "none")

    def GetDevicePluginOptions(self, request, context):
        return pb.DevicePluginOptions(get_preferred_allocation_available=self.policy != "none")

    def _device_list(self) -> "pb.ListAndWatchResponse":
        resp = pb.ListAndWatchResponse()
        for dev in self.ngm.list_devices().values():
            resp.devices.append(_device_msg(dev))
        return resp

    def ListAndWatch(self, request, context):
        log.info("device-plugin: ListAndWatch start")
        yield self._device_list()
        while context.is_active() and self.ngm.grpc_server is not None:
            try:
                d = self.ngm.health.get(timeout=0.5)
            except queue.Empty:
                continue
            log.info("device-plugin: %s device marked as %s", d.id, d.health)
            self.ngm.set_device_health(d.id, d.health, d.numa_node)
            yield self._device_list()

    def Allocate(self, request, context):
        resps = pb.AllocateResponse()
        strategy = self.ngm.gpu_config.sharing.strategy
        for rqt in request.container_requests:
            ids = list(rqt.devices_ids)
            try:
                sharing.validate_request(ids, len(self.ngm.list_physical_devices()), strategy)
                resp = pb.ContainerAllocateResponse()
                for device_id in ids:
                    for spec in self.ngm.device_spec(device_id):
                        resp.devices.add(host_path=spec.host_path, container_path=spec.container_path, permissions=spec.permissions)
            except (ValueError, RuntimeError) as e:    # SharingError / AllocationError
                context.abort(grpc.StatusCode.UNKNOWN, str(e))
            for d in self.ngm.default_devices:
                resp.devices.add(host_path=d, container_path=d, permissions="mrw")
            for m in self.ngm.mount_paths:
                resp.mounts.add(host_path=m.host_path, container_path=m.container_path, read_only=m.read_only)
            for k, v in self.ngm.envs(len(ids)).items():
                resp.envs[k] = v
            transport.apply(self.ngm.gpu_config.transport, self.ngm.mount_paths, resp)
            resps.container_responses.append(resp)
        return resps

    def PreStartContainer(self, request, context):
        log.error("device-plugin: PreStart should NOT be called for the B200 GPU device plugin")
        return pb.PreStartContainerResponse()

    def GetPreferredAllocation(self, request, context):
        resp = pb.PreferredAllocationResponse()
        if self.policy == "none":
            log.error("device-plugin: GetPreferredAllocation should NOT be called for the B200 GPU device plugin")
            return resp
        devices = self.ngm.list_devices()
        numa_of = lambda d: devices[d].numa_node if d in devices else None
        for rqt in request.container_requests:
            ids = preferred.preferred_allocation(list(rqt.available_deviceIDs), list(rqt.must_include_deviceIDs), rqt.allocation_size, numa_of, self.policy)
            resp.container_responses.add(deviceIDs=ids)
        return resp

    # ------------------------------------------------------------------ wiring
    def register(self, server: grpc.Server) -> None:
        uu = grpc.unary_unary_rpc_method_handler
        handlers = {
            "GetDevicePluginOptions": uu(self.GetDevicePluginOptions, pb.Empty.FromString, pb.DevicePluginOptions.SerializeToString),
            "ListAndWatch": grpc.unary_stream_rpc_method_handler(self.ListAndWatch, pb.Empty.FromString, pb.ListAndWatchResponse.SerializeToString),
            "Allocate": uu(self.Allocate, pb.AllocateRequest.FromString, pb.AllocateResponse.SerializeToString),
            "PreStartContainer": uu(self.PreStartContainer, pb.PreStartContainerRequest.FromString, pb.PreStartContainerResponse.SerializeToString),
            "GetPreferredAllocation": uu(self.GetPreferredAllocation, pb.PreferredAllocationRequest.FromString, pb.PreferredAllocationResponse.SerializeToString),
        }
        server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(protos.DEVICE_PLUGIN_SERVICE, handlers),))


def register_with_kubelet(kubelet_socket: str, plugin_endpoint: str, resource_name: str, timeout: float = 10.0, preferred_allocation: bool = False) -> None:
    with grpc.insecure_channel(f"unix:{kubelet_socket}") as ch:
        grpc.channel_ready_future(ch).result(timeout=timeout)
        call = ch.unary_unary(f"/{protos.REGISTRATION_SERVICE}/Register", request_serializer=pb.RegisterRequest.SerializeToString,
                              response_deserializer=pb.Empty.FromString)
        req = pb.RegisterRequest(version=protos.DEVICE_PLUGIN_VERSION, endpoint=plugin_endpoint, resource_name=resource_name)
        if preferred_allocation:            # the reference never sets options; only the opt-in policy does
            req.options.get_preferred_allocation_available = True
        call(req, timeout=timeout)


class DevicePluginClient:
    """What the kubelet does to a plugin; used by the conformance harness and tests."""

    def __init__(self, socket_path: str):
        self.channel = grpc.insecure_channel(f"unix:{socket_path}")
        svc = protos.DEVICE_PLUGIN_SERVICE
        self._options = self.channel.unary_unary(f"/{svc}/GetDevicePluginOptions", request_serializer=pb.Empty.SerializeToString, response_deserializer=pb.DevicePluginOptions.FromString)
        self._law = self.channel.unary_stream(f"/{svc}/ListAndWatch", request_serializer=pb.Empty.SerializeToString, response_deserializer=pb.ListAndWatchResponse.FromString)
        self._alloc = self.channel.unary_unary(f"/{svc}/Allocate", request_serializer=pb.AllocateRequest.SerializeToString, response_deserializer=pb.AllocateResponse.FromString)
        self._pref = self.channel.unary_unary(f"/{svc}/GetPreferredAllocation", request_serializer=pb.PreferredAllocationRequest.SerializeToString,
                                              response_deserializer=pb.PreferredAllocationResponse.FromString)

    def wait_ready(self, timeout: float = 10.0) -> None:
        grpc.channel_ready_future(self.channel).result(timeout=timeout)

    def options(self):
        return self._options(pb.Empty(), timeout=5)

    def list_and_watch(self):
        return self._law(pb.Empty())

    def allocate(self, *container_device_ids: list):
        req = pb.AllocateRequest()
        for ids in container_device_ids:
            req.container_requests.add(devices_ids=list(ids))
        return self._alloc(req, timeout=5)

    def preferred(self, available: list, must_include: list, size: int) -> list:
        req = pb.PreferredAllocationRequest()
        req.container_requests.add(available_deviceIDs=list(available), must_include_deviceIDs=list(must_include), allocation_size=size)
        resps = self._pref(req, timeout=5).container_responses
        return list(resps[0].deviceIDs) if resps else []

    def close(self) -> None:
        self.channel.close()
