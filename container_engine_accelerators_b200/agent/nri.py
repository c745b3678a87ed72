"""NRI device injector: a containerd NRI plugin that injects device nodes named in a pod annotation.

Contract: reference nri_device_injector/nri_device_injector.go:30-199: plugin name
`device_injector_nri`, index `10`, socket /var/run/nri/nri.sock; annotation `devices.gke.io/container.<name>` is a YAML
list of {path,type,major,minor,file_mode,uid,gid} of which only `path` and non-zero file_mode/uid/gid are honoured —
type/major/minor always come from lstat; duplicate paths: first wins; bad YAML or an un-stat-able path fails the
container creation; pod == nil is a no-op.

The wire stack is written out here because nothing in the image speaks it: NRI multiplexes two ttrpc connections over
one Unix socket (8-byte frames: conn id + length, big endian; conn 1 = Plugin service served by us, conn 2 = Runtime
service we call), and ttrpc frames are a 10-byte header (length, stream id, type, flags) + a protobuf Request/Response
(reference: vendor/github.com/containerd/nri/pkg/net/multiplex/mux.go:139-141,243-244,
vendor/github.com/containerd/ttrpc/channel.go:32-60, vendor/github.com/containerd/ttrpc/request.proto:9-28).
"""
from __future__ import annotations

import logging
import os
import queue
import socket
import stat as statmod
import struct
import threading
from typing import Callable, Optional

import yaml

from . import protos
from .protos import nri as pb

log = logging.getLogger("b200-nri-device-injector")

DEVICE_KEY_PREFIX = "devices.gke.io"
CTR_DEVICE_KEY_PREFIX = DEVICE_KEY_PREFIX + "/container."
PLUGIN_NAME = "device_injector_nri"
PLUGIN_IDX = "10"
DEFAULT_SOCKET = "/var/run/nri/nri.sock"
PLUGIN_SERVICE_CONN, RUNTIME_SERVICE_CONN = 1, 2
MSG_REQUEST, MSG_RESPONSE = 1, 2
MAX_PAYLOAD = 1 << 24


# --------------------------------------------------------------------------------------------- device logic
class DeviceError(ValueError):
    pass


def get_devices(ctr_name: str, annotations: dict) -> list:
    """Parsed, de-duplicated (first wins) device dicts for one container, [] when not annotated."""
    value = (annotations or {}).get(CTR_DEVICE_KEY_PREFIX + ctr_name)
    if value is None:
        return []
    try:
        parsed = yaml.safe_load(value)
    except yaml.YAMLError as e:
        raise DeviceError(f"invalid device annotation \"{CTR_DEVICE_KEY_PREFIX + ctr_name}\": {e}") from e
    if parsed is None:
        return []
    if not isinstance(parsed, list) or not all(isinstance(d, dict) for d in parsed):
        raise DeviceError(f"invalid device annotation \"{CTR_DEVICE_KEY_PREFIX + ctr_name}\": expected a list of device objects")
    seen, out = set(), []
    for d in parsed:
        path = d.get("path", "")
        if not isinstance(path, str):
            raise DeviceError(f"invalid device annotation \"{CTR_DEVICE_KEY_PREFIX + ctr_name}\": path must be a string")
        if path in seen:
            continue
        seen.add(path)
        out.append(d)
    return out


def to_nri_device(dev: dict, lstat: Callable = os.lstat):
    path = dev.get("path", "")
    try:
        st = lstat(path)
    except OSError as e:
        raise DeviceError(f"failed to get info from device path {path}: {e}") from e
    fmt = statmod.S_IFMT(st.st_mode)
    if fmt == statmod.S_IFBLK:
        dev_type = "b"
    elif fmt == statmod.S_IFCHR:
        dev_type = "c"
    elif fmt == statmod.S_IFIFO:
        dev_type = "p"
    else:
        raise DeviceError(f"invalid device type {st.st_mode} from device path {path}")
    out = pb.LinuxDevice(path=path, type=dev_type, major=os.major(st.st_rdev), minor=os.minor(st.st_rdev))
    if int(dev.get("file_mode") or 0):
        out.file_mode.value = int(dev["file_mode"])
    if int(dev.get("uid") or 0):
        out.uid.value = int(dev["uid"])
    if int(dev.get("gid") or 0):
        out.gid.value = int(dev["gid"])
    return out


def create_container(pod, container, lstat: Callable = os.lstat):
    """The CreateContainer hook: returns a ContainerAdjustment (possibly empty) or raises DeviceError."""
    adjust = pb.ContainerAdjustment()
    if pod is None:
        return adjust
    for d in get_devices(container.name, dict(pod.annotations)):
        log.info("Annotated device %s (container=%s pod=%s/%s)", d.get("path"), container.name, pod.namespace, pod.name)
        adjust.linux.devices.append(to_nri_device(d, lstat))
    return adjust


# --------------------------------------------------------------------------------------------- wire: mux
class Mux:
    """Two logical byte streams over one socket."""

    def __init__(self, sock: socket.socket):
        self.sock = sock
        self.wlock = threading.Lock()
        self.queues = {PLUGIN_SERVICE_CONN: queue.Queue(), RUNTIME_SERVICE_CONN: queue.Queue()}
        self.closed = threading.Event()
        self.reader = threading.Thread(target=self._read_loop, daemon=True)
        self.reader.start()

    def _recv_exact(self, n: int) -> Optional[bytes]:
        buf = bytearray()
        while len(buf) < n:
            try:
                chunk = self.sock.recv(n - len(buf))
            except OSError:
                return None
            if not chunk:
                return None
            buf += chunk
        return bytes(buf)

    def _read_loop(self) -> None:
        while True:
            hdr = self._recv_exact(8)
            if hdr is None:
                break
            cid, cnt = struct.unpack(">II", hdr)
            if cnt > MAX_PAYLOAD:
                break
            payload = self._recv_exact(cnt) if cnt else b""
            if payload is None:
                break
            q = self.queues.get(cid)
            if q is not None:
                q.put(payload)
        self.closed.set()
        for q in self.queues.values():
            q.put(None)

    def write(self, cid: int, data: bytes) -> None:
        with self.wlock:
            self.sock.sendall(struct.pack(">II", cid, len(data)) + data)

    def close(self) -> None:
        try:
            self.sock.shutdown(socket.SHUT_RDWR)
        except OSError:
            pass
        self.sock.close()


class MuxStream:
    """Blocking byte-stream view of one mux connection."""

    def __init__(self, mux: Mux, cid: int):
        self.mux, self.cid, self.buf = mux, cid, bytearray()

    def read_exact(self, n: int) -> Optional[bytes]:
        while len(self.buf) < n:
            chunk = self.mux.queues[self.cid].get()
            if chunk is None:
                return None
            self.buf += chunk
        out = bytes(self.buf[:n])
        del self.buf[:n]
        return out

    def write(self, data: bytes) -> None:
        self.mux.write(self.cid, data)


# --------------------------------------------------------------------------------------------- wire: ttrpc
def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _field_bytes(num: int, data: bytes) -> bytes:
    return _varint((num << 3) | 2) + _varint(len(data)) + data


def _parse_fields(data: bytes) -> dict:
    out, i = {}, 0
    while i < len(data):
        key, shift = 0, 0
        while True:
            b = data[i]; i += 1
            key |= (b & 0x7F) << shift; shift += 7
            if not b & 0x80:
                break
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, shift = 0, 0
            while True:
                b = data[i]; i += 1
                val |= (b & 0x7F) << shift; shift += 7
                if not b & 0x80:
                    break
            out.setdefault(num, []).append(val)
        elif wt == 2:
            ln, shift = 0, 0
            while True:
                b = data[i]; i += 1
                ln |= (b & 0x7F) << shift; shift += 7
                if not b & 0x80:
                    break
            out.setdefault(num, []).append(data[i:i + ln]); i += ln
        elif wt == 1:
            out.setdefault(num, []).append(data[i:i + 8]); i += 8
        elif wt == 5:
            out.setdefault(num, []).append(data[i:i + 4]); i += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
    return out


def encode_request(service: str, method: str, payload: bytes) -> bytes:
    return _field_bytes(1, service.encode()) + _field_bytes(2, method.encode()) + _field_bytes(3, payload)


def decode_request(data: bytes):
    f = _parse_fields(data)
    return (f.get(1, [b""])[0].decode(), f.get(2, [b""])[0].decode(), f.get(3, [b""])[0])


def encode_response(payload: bytes, code: int = 0, message: str = "") -> bytes:
    out = b""
    if code:
        status = _varint((1 << 3) | 0) + _varint(code) + _field_bytes(2, message.encode())
        out += _field_bytes(1, status)
    return out + _field_bytes(2, payload)


def decode_response(data: bytes):
    f = _parse_fields(data)
    code, msg = 0, ""
    if 1 in f:
        s = _parse_fields(f[1][0])
        code = s.get(1, [0])[0]
        msg = s.get(2, [b""])[0].decode(errors="replace")
    return code, msg, f.get(2, [b""])[0]


def write_message(stream: MuxStream, stream_id: int, mtype: int, payload: bytes) -> None:
    stream.write(struct.pack(">IIBB", len(payload), stream_id, mtype, 0) + payload)


def read_message(stream: MuxStream):
    hdr = stream.read_exact(10)
    if hdr is None:
        return None
    length, sid, mtype, flags = struct.unpack(">IIBB", hdr)
    payload = stream.read_exact(length) if length else b""
    if payload is None:
        return None
    return sid, mtype, flags, payload


class TtrpcClient:
    def __init__(self, stream: MuxStream):
        self.stream, self.next_id, self.lock = stream, 1, threading.Lock()

    def call(self, service: str, method: str, request, response_cls):
        with self.lock:
            sid = self.next_id
            self.next_id += 2
            write_message(self.stream, sid, MSG_REQUEST, encode_request(service, method, request.SerializeToString()))
            while True:
                msg = read_message(self.stream)
                if msg is None:
                    raise ConnectionError("ttrpc connection closed")
                rsid, mtype, _, payload = msg
                if mtype == MSG_RESPONSE and rsid == sid:
                    code, text, body = decode_response(payload)
                    if code:
                        raise RuntimeError(f"{service}/{method}: status {code}: {text}")
                    return response_cls.FromString(body)


class TtrpcServer:
    """Serves {method: (request_cls, handler)} for one service on a stream until the stream closes."""

    def __init__(self, stream: MuxStream, service: str, methods: dict):
        self.stream, self.service, self.methods = stream, service, methods

    def serve(self) -> None:
        while True:
            msg = read_message(self.stream)
            if msg is None:
                return
            sid, mtype, _, payload = msg
            if mtype != MSG_REQUEST:
                continue
            service, method, body = decode_request(payload)
            entry = self.methods.get(method) if service == self.service else None
            if entry is None:
                write_message(self.stream, sid, MSG_RESPONSE, encode_response(b"", 12, f"unimplemented {service}/{method}"))
                continue
            req_cls, handler = entry
            try:
                resp = handler(req_cls.FromString(body))
                write_message(self.stream, sid, MSG_RESPONSE, encode_response(resp.SerializeToString()))
            except Exception as e:    # an error fails the runtime's operation (container creation), as in the reference
                log.warning("%s failed: %s", method, e)
                write_message(self.stream, sid, MSG_RESPONSE, encode_response(b"", 2, str(e)))


# --------------------------------------------------------------------------------------------- the plugin
class DeviceInjectorPlugin:
    def __init__(self, socket_path: str = DEFAULT_SOCKET, name: str = PLUGIN_NAME, idx: str = PLUGIN_IDX, lstat: Callable = os.lstat):
        self.socket_path, self.name, self.idx, self.lstat = socket_path, name, idx, lstat
        self.mux: Optional[Mux] = None
        self.configured = threading.Event()

    # Plugin service handlers
    def _configure(self, req):
        log.info("configured by runtime %s %s", req.runtime_name, req.runtime_version)
        self.configured.set()
        return pb.ConfigureResponse(events=protos.NRI_EVENT_CREATE_CONTAINER)

    def _synchronize(self, req):
        return pb.SynchronizeResponse()

    def _create_container(self, req):
        pod = req.pod if req.HasField("pod") else None
        log.info("Started CreateContainer container=%s", req.container.name)
        adjust = create_container(pod, req.container, self.lstat)
        log.info("Finished CreateContainer container=%s (%d devices)", req.container.name, len(adjust.linux.devices))
        return pb.CreateContainerResponse(adjust=adjust)

    def _noop(self, req):
        return pb.Empty()

    def run(self) -> None:
        """Connect, register, serve until the runtime closes the connection (then return: the pod restarts us)."""
        sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        sock.connect(self.socket_path)
        self.mux = Mux(sock)
        server = TtrpcServer(MuxStream(self.mux, PLUGIN_SERVICE_CONN), protos.NRI_PLUGIN_SERVICE, {
            "Configure": (pb.ConfigureRequest, self._configure),
            "Synchronize": (pb.SynchronizeRequest, self._synchronize),
            "CreateContainer": (pb.CreateContainerRequest, self._create_container),
            "StateChange": (pb.StateChangeEvent, self._noop),
            "Shutdown": (pb.Empty, self._noop),
        })
        t = threading.Thread(target=server.serve, daemon=True)
        t.start()
        client = TtrpcClient(MuxStream(self.mux, RUNTIME_SERVICE_CONN))
        client.call(protos.NRI_RUNTIME_SERVICE, "RegisterPlugin", pb.RegisterPluginRequest(plugin_name=self.name, plugin_idx=self.idx), pb.Empty)
        log.info("registered NRI plugin %s-%s", self.idx, self.name)
        t.join()
        log.info("NRI connection closed")

    def close(self) -> None:
        if self.mux:
            self.mux.close()


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(prog="b200-nri-device-injector")
    ap.add_argument("--socket", default=DEFAULT_SOCKET)
    ap.add_argument("--name", default=PLUGIN_NAME)
    ap.add_argument("--idx", default=PLUGIN_IDX)
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname).1s %(name)s] %(message)s")
    try:
        DeviceInjectorPlugin(args.socket, args.name, args.idx).run()
    except Exception as e:
        log.error("plugin exited with error %s", e)
        return 1
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main())
