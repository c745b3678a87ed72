// Hand-rolled protobuf wire codec for the dozen small messages of the kubelet deviceplugin v1beta1 and
// podresources v1alpha1 APIs (field numbers: reference vendor/k8s.io/kubelet/pkg/apis/deviceplugin/v1beta1/api.proto:24-211).
// Only varint (0) and length-delimited (2) wire types occur in those messages.
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

namespace pb {

inline void put_varint(std::string* out, uint64_t v) {
  while (v >= 128) { out->push_back((char)(0x80 | (v & 0x7F))); v >>= 7; }
  out->push_back((char)v);
}
inline void put_tag(std::string* out, int field, int wt) { put_varint(out, ((uint64_t)field << 3) | (uint64_t)wt); }
inline void put_bytes(std::string* out, int field, const std::string& s) { put_tag(out, field, 2); put_varint(out, s.size()); *out += s; }
inline void put_string(std::string* out, int field, const std::string& s) { if (!s.empty()) put_bytes(out, field, s); }
inline void put_bool(std::string* out, int field, bool b) { if (b) { put_tag(out, field, 0); put_varint(out, 1); } }
inline void put_int(std::string* out, int field, int64_t v) { if (v) { put_tag(out, field, 0); put_varint(out, (uint64_t)v); } }
inline void put_map_entry(std::string* out, int field, const std::string& k, const std::string& v) {
  std::string e;
  put_bytes(&e, 1, k); put_bytes(&e, 2, v);
  put_bytes(out, field, e);
}

struct Field { int number; int wire_type; uint64_t varint; std::string bytes; };

// Returns false on malformed input. Unknown wire types 1/5 are skipped (never produced by these APIs, but be lenient).
inline bool parse(const std::string& in, std::vector<Field>* out) {
  size_t i = 0;
  auto varint = [&](uint64_t* v) {
    *v = 0; int shift = 0;
    while (i < in.size()) { uint8_t b = (uint8_t)in[i++]; *v |= (uint64_t)(b & 0x7F) << shift; if (!(b & 0x80)) return true; shift += 7; if (shift > 63) return false; }
    return false;
  };
  while (i < in.size()) {
    uint64_t key; if (!varint(&key)) return false;
    Field f{(int)(key >> 3), (int)(key & 7), 0, ""};
    if (f.wire_type == 0) { if (!varint(&f.varint)) return false; }
    else if (f.wire_type == 2) { uint64_t n; if (!varint(&n) || i + n > in.size()) return false; f.bytes = in.substr(i, (size_t)n); i += (size_t)n; }
    else if (f.wire_type == 1) { if (i + 8 > in.size()) return false; i += 8; }
    else if (f.wire_type == 5) { if (i + 4 > in.size()) return false; i += 4; }
    else return false;
    out->push_back(std::move(f));
  }
  return true;
}

// ---- deviceplugin v1beta1
struct Device { std::string id, health; bool has_numa = false; int64_t numa = 0; };
struct DeviceSpec { std::string container_path, host_path, permissions; };
struct Mount { std::string container_path, host_path; bool read_only = false; };
struct ContainerAllocateResponse { std::map<std::string, std::string> envs; std::vector<Mount> mounts; std::vector<DeviceSpec> devices; };

inline std::string encode_device(const Device& d) {
  std::string o;
  put_string(&o, 1, d.id); put_string(&o, 2, d.health);
  if (d.has_numa) { std::string node; put_tag(&node, 1, 0); put_varint(&node, (uint64_t)d.numa); std::string topo; put_bytes(&topo, 1, node); put_bytes(&o, 3, topo); }
  return o;
}
inline std::string encode_list_and_watch(const std::vector<Device>& devs) { std::string o; for (auto& d : devs) put_bytes(&o, 1, encode_device(d)); return o; }
// DevicePluginOptions{pre_start_required = 1, get_preferred_allocation_available = 2}
inline std::string encode_options(bool preferred_allocation) { std::string o; put_bool(&o, 2, preferred_allocation); return o; }
inline std::string encode_register_request(const std::string& version, const std::string& endpoint, const std::string& resource, bool preferred_allocation = false) {
  std::string o; put_string(&o, 1, version); put_string(&o, 2, endpoint); put_string(&o, 3, resource);
  if (preferred_allocation) put_bytes(&o, 4, encode_options(true));      // the reference never sends options (field absent)
  return o;
}
// PreferredAllocationRequest{repeated ContainerPreferredAllocationRequest{repeated available_deviceIDs = 1, repeated must_include_deviceIDs = 2, int32 allocation_size = 3} = 1}
struct PreferredRequest { std::vector<std::string> available, must_include; int64_t size = 0; };
inline bool decode_preferred_request(const std::string& in, std::vector<PreferredRequest>* out) {
  std::vector<Field> fs;
  if (!parse(in, &fs)) return false;
  for (auto& f : fs) if (f.number == 1 && f.wire_type == 2) {
    std::vector<Field> cf; if (!parse(f.bytes, &cf)) return false;
    PreferredRequest r;
    for (auto& x : cf) {
      if (x.number == 1 && x.wire_type == 2) r.available.push_back(x.bytes);
      if (x.number == 2 && x.wire_type == 2) r.must_include.push_back(x.bytes);
      if (x.number == 3 && x.wire_type == 0) r.size = (int64_t)x.varint;
    }
    out->push_back(r);
  }
  return true;
}
// PreferredAllocationResponse{repeated ContainerPreferredAllocationResponse{repeated deviceIDs = 1} = 1}
inline std::string encode_preferred_response(const std::vector<std::vector<std::string>>& rs) {
  std::string o;
  for (auto& ids : rs) { std::string c; for (auto& id : ids) put_bytes(&c, 1, id); put_bytes(&o, 1, c); }
  return o;
}
inline std::string encode_allocate_response(const std::vector<ContainerAllocateResponse>& rs) {
  std::string o;
  for (auto& r : rs) {
    std::string c;
    for (auto& kv : r.envs) put_map_entry(&c, 1, kv.first, kv.second);
    for (auto& m : r.mounts) { std::string s; put_string(&s, 1, m.container_path); put_string(&s, 2, m.host_path); put_bool(&s, 3, m.read_only); put_bytes(&c, 2, s); }
    for (auto& d : r.devices) { std::string s; put_string(&s, 1, d.container_path); put_string(&s, 2, d.host_path); put_string(&s, 3, d.permissions); put_bytes(&c, 3, s); }
    put_bytes(&o, 1, c);
  }
  return o;
}
// AllocateRequest{repeated ContainerAllocateRequest{repeated string devices_ids = 1} = 1}
inline bool decode_allocate_request(const std::string& in, std::vector<std::vector<std::string>>* out) {
  std::vector<Field> fs;
  if (!parse(in, &fs)) return false;
  for (auto& f : fs) if (f.number == 1 && f.wire_type == 2) {
    std::vector<Field> cf; if (!parse(f.bytes, &cf)) return false;
    std::vector<std::string> ids;
    for (auto& x : cf) if (x.number == 1 && x.wire_type == 2) ids.push_back(x.bytes);
    out->push_back(ids);
  }
  return true;
}

// ---- podresources v1alpha1: ListPodResourcesResponse{repeated PodResources{name=1, namespace=2, repeated ContainerResources{name=1, repeated ContainerDevices{resource_name=1, repeated device_ids=2}=2}=3}=1}
struct ContainerDevices { std::string ns, pod, container, resource; std::vector<std::string> ids; };
inline bool decode_pod_resources(const std::string& in, std::vector<ContainerDevices>* out) {
  std::vector<Field> top; if (!parse(in, &top)) return false;
  for (auto& p : top) if (p.number == 1 && p.wire_type == 2) {
    std::vector<Field> pf; if (!parse(p.bytes, &pf)) return false;
    std::string name, ns;
    for (auto& f : pf) { if (f.number == 1) name = f.bytes; if (f.number == 2) ns = f.bytes; }
    for (auto& f : pf) if (f.number == 3 && f.wire_type == 2) {
      std::vector<Field> cf; if (!parse(f.bytes, &cf)) return false;
      std::string cname;
      for (auto& x : cf) if (x.number == 1) cname = x.bytes;
      for (auto& x : cf) if (x.number == 2 && x.wire_type == 2) {
        std::vector<Field> df; if (!parse(x.bytes, &df)) return false;
        ContainerDevices cd{ns, name, cname, "", {}};
        for (auto& y : df) { if (y.number == 1) cd.resource = y.bytes; if (y.number == 2) cd.ids.push_back(y.bytes); }
        out->push_back(cd);
      }
    }
  }
  return true;
}

}  // namespace pb
