"""Runtime-built protobuf descriptors for the wire protocols the node agent speaks.

No protoc / grpc_tools in the image, so the message classes are built from
FileDescriptorProto objects at import time. Field numbers and names follow the public Kubernetes
APIs the reference vendors:
  * deviceplugin v1beta1  (reference: vendor/k8s.io/kubelet/pkg/apis/deviceplugin/v1beta1/api.proto:24-211)
  * podresources v1alpha1 (reference: pkg/gpu/nvidia/metrics/devices.go:33-34,69-95)
  * the NRI plugin API subset used by the device injector (reference: nri_device_injector/nri_device_injector.go:86-123)
"""
from __future__ import annotations

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_T = descriptor_pb2.FieldDescriptorProto
_SCALAR = {
    "string": _T.TYPE_STRING, "bool": _T.TYPE_BOOL, "int32": _T.TYPE_INT32, "int64": _T.TYPE_INT64,
    "uint32": _T.TYPE_UINT32, "uint64": _T.TYPE_UINT64, "bytes": _T.TYPE_BYTES,
}


def _build(package: str, filename: str, messages: dict) -> dict:
    """messages: {Name: [(field_name, number, type, repeated?)]}; type may be a scalar name, a message
    name from the same file, or ("map", key_type, value_type)."""
    fd = descriptor_pb2.FileDescriptorProto(name=filename, package=package, syntax="proto3")
    for mname, fields in messages.items():
        m = fd.message_type.add(name=mname)
        for f in fields:
            fname, num, ftype = f[0], f[1], f[2]
            repeated = len(f) > 3 and f[3]
            fld = m.field.add(name=fname, number=num)
            if isinstance(ftype, tuple) and ftype[0] == "map":
                entry = m.nested_type.add(name="".join(p.capitalize() for p in fname.split("_")) + "Entry")
                entry.options.map_entry = True
                k = entry.field.add(name="key", number=1, label=_T.LABEL_OPTIONAL, type=_SCALAR[ftype[1]])
                v = entry.field.add(name="value", number=2, label=_T.LABEL_OPTIONAL)
                if ftype[2] in _SCALAR:
                    v.type = _SCALAR[ftype[2]]
                else:
                    v.type = _T.TYPE_MESSAGE
                    v.type_name = f".{package}.{ftype[2]}"
                del k
                fld.label = _T.LABEL_REPEATED
                fld.type = _T.TYPE_MESSAGE
                fld.type_name = f".{package}.{mname}.{entry.name}"
            else:
                fld.label = _T.LABEL_REPEATED if repeated else _T.LABEL_OPTIONAL
                if ftype in _SCALAR:
                    fld.type = _SCALAR[ftype]
                else:
                    fld.type = _T.TYPE_MESSAGE
                    fld.type_name = f".{package}.{ftype}"
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    out = {}
    for mname in messages:
        out[mname] = message_factory.GetMessageClass(pool.FindMessageTypeByName(f"{package}.{mname}"))
    return out


class _NS:
    def __init__(self, d: dict):
        self.__dict__.update(d)


# --------------------------------------------------------------------------- deviceplugin v1beta1
deviceplugin = _NS(_build("v1beta1", "b200/deviceplugin_v1beta1.proto", {
    "DevicePluginOptions": [("pre_start_required", 1, "bool"), ("get_preferred_allocation_available", 2, "bool")],
    "RegisterRequest": [("version", 1, "string"), ("endpoint", 2, "string"), ("resource_name", 3, "string"),
                        ("options", 4, "DevicePluginOptions")],
    "Empty": [],
    "ListAndWatchResponse": [("devices", 1, "Device", True)],
    "TopologyInfo": [("nodes", 1, "NUMANode", True)],
    "NUMANode": [("ID", 1, "int64")],
    "Device": [("ID", 1, "string"), ("health", 2, "string"), ("topology", 3, "TopologyInfo")],
    "PreStartContainerRequest": [("devices_ids", 1, "string", True)],
    "PreStartContainerResponse": [],
    "PreferredAllocationRequest": [("container_requests", 1, "ContainerPreferredAllocationRequest", True)],
    "ContainerPreferredAllocationRequest": [("available_deviceIDs", 1, "string", True), ("must_include_deviceIDs", 2, "string", True),
                                            ("allocation_size", 3, "int32")],
    "PreferredAllocationResponse": [("container_responses", 1, "ContainerPreferredAllocationResponse", True)],
    "ContainerPreferredAllocationResponse": [("deviceIDs", 1, "string", True)],
    "AllocateRequest": [("container_requests", 1, "ContainerAllocateRequest", True)],
    "ContainerAllocateRequest": [("devices_ids", 1, "string", True)],
    "AllocateResponse": [("container_responses", 1, "ContainerAllocateResponse", True)],
    "ContainerAllocateResponse": [("envs", 1, ("map", "string", "string")), ("mounts", 2, "Mount", True),
                                  ("devices", 3, "DeviceSpec", True), ("annotations", 4, ("map", "string", "string"))],
    "Mount": [("container_path", 1, "string"), ("host_path", 2, "string"), ("read_only", 3, "bool")],
    "DeviceSpec": [("container_path", 1, "string"), ("host_path", 2, "string"), ("permissions", 3, "string")],
}))

DEVICE_PLUGIN_VERSION = "v1beta1"
HEALTHY = "Healthy"
UNHEALTHY = "Unhealthy"
REGISTRATION_SERVICE = "v1beta1.Registration"
DEVICE_PLUGIN_SERVICE = "v1beta1.DevicePlugin"

# --------------------------------------------------------------------------- podresources v1alpha1
podresources = _NS(_build("v1alpha1", "b200/podresources_v1alpha1.proto", {
    "ListPodResourcesRequest": [],
    "ListPodResourcesResponse": [("pod_resources", 1, "PodResources", True)],
    "PodResources": [("name", 1, "string"), ("namespace", 2, "string"), ("containers", 3, "ContainerResources", True)],
    "ContainerResources": [("name", 1, "string"), ("devices", 2, "ContainerDevices", True)],
    "ContainerDevices": [("resource_name", 1, "string"), ("device_ids", 2, "string", True)],
}))
POD_RESOURCES_SERVICE = "v1alpha1.PodResourcesLister"
POD_RESOURCES_SERVICE_V1 = "v1.PodResourcesLister"      # same List() request/response field numbers for what we read (name, namespace, containers.devices)

# --------------------------------------------------------------------------- NRI (subset)
# Field numbers follow containerd/nri pkg/api/api.proto (v0.5): only what a CreateContainer-time
# device injector touches. Unknown fields from the runtime are preserved by protobuf and ignored.
nri = _NS(_build("nri.pkg.api.v1alpha1", "b200/nri_api.proto", {
    "RegisterPluginRequest": [("plugin_name", 1, "string"), ("plugin_idx", 2, "string")],
    "Empty": [],
    "ConfigureRequest": [("config", 1, "string"), ("runtime_name", 2, "string"), ("runtime_version", 3, "string")],
    "ConfigureResponse": [("events", 2, "int32")],
    "SynchronizeRequest": [("pods", 1, "PodSandbox", True), ("containers", 2, "Container", True)],
    "SynchronizeResponse": [("update", 1, "ContainerUpdate", True)],
    "PodSandbox": [("id", 1, "string"), ("name", 2, "string"), ("uid", 3, "string"), ("namespace", 4, "string"),
                   ("labels", 5, ("map", "string", "string")), ("annotations", 6, ("map", "string", "string"))],
    "Container": [("id", 1, "string"), ("pod_sandbox_id", 2, "string"), ("name", 3, "string")],
    "CreateContainerRequest": [("pod", 1, "PodSandbox"), ("container", 2, "Container")],
    "CreateContainerResponse": [("adjust", 1, "ContainerAdjustment"), ("update", 2, "ContainerUpdate", True)],
    "ContainerUpdate": [("container_id", 1, "string")],
    "ContainerAdjustment": [("annotations", 2, ("map", "string", "string")), ("linux", 6, "LinuxContainerAdjustment")],
    "LinuxContainerAdjustment": [("devices", 1, "LinuxDevice", True)],
    "LinuxDevice": [("path", 1, "string"), ("type", 2, "string"), ("major", 3, "int64"), ("minor", 4, "int64"),
                    ("file_mode", 5, "OptionalFileMode"), ("uid", 6, "OptionalUInt32"), ("gid", 7, "OptionalUInt32")],
    "OptionalFileMode": [("value", 1, "uint32")],
    "OptionalUInt32": [("value", 1, "uint32")],
    "StateChangeEvent": [("event", 1, "int32"), ("pod", 2, "PodSandbox"), ("container", 3, "Container")],
}))
NRI_PLUGIN_SERVICE = "nri.pkg.api.v1alpha1.Plugin"
NRI_RUNTIME_SERVICE = "nri.pkg.api.v1alpha1.Runtime"
NRI_EVENT_CREATE_CONTAINER = 1 << 3   # Event_CREATE_CONTAINER = 4 -> mask bit (event-1)
