// The slice of the NVML C ABI the node agent uses, declared here so the native tools build with nothing but a C++
// compiler (no CUDA toolkit, no nvml.h in the build image). libnvidia-ml.so.1 is resolved with dlopen at run time
// (b200agent_nvml.cc), exactly as the reference's go-nvml binding does (vendor/github.com/NVIDIA/go-nvml/pkg/nvml/lib.go:30-31).
// Only facts of the binary interface are restated: enumerator values, struct layouts and entry-point signatures of
// NVML 12.x; `static_assert`s pin the layouts. If <nvml.h> was included first, this header stands aside.
#pragma once
#ifndef __nvml_nvml_h__

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nvmlDevice_st* nvmlDevice_t;        // opaque handles
typedef struct nvmlEventSet_st* nvmlEventSet_t;

typedef enum nvmlReturn_enum {
  NVML_SUCCESS = 0,
  NVML_ERROR_UNINITIALIZED = 1,
  NVML_ERROR_INVALID_ARGUMENT = 2,
  NVML_ERROR_NOT_SUPPORTED = 3,
  NVML_ERROR_NO_PERMISSION = 4,
  NVML_ERROR_NOT_FOUND = 6,
  NVML_ERROR_INSUFFICIENT_SIZE = 7,
  NVML_ERROR_DRIVER_NOT_LOADED = 9,
  NVML_ERROR_TIMEOUT = 10,
  NVML_ERROR_GPU_IS_LOST = 15,
  NVML_ERROR_UNKNOWN = 999
} nvmlReturn_t;

typedef enum nvmlSamplingType_enum { NVML_TOTAL_POWER_SAMPLES = 0, NVML_GPU_UTILIZATION_SAMPLES = 1, NVML_MEMORY_UTILIZATION_SAMPLES = 2 } nvmlSamplingType_t;
typedef enum nvmlValueType_enum { NVML_VALUE_TYPE_DOUBLE = 0, NVML_VALUE_TYPE_UNSIGNED_INT = 1, NVML_VALUE_TYPE_UNSIGNED_LONG = 2, NVML_VALUE_TYPE_UNSIGNED_LONG_LONG = 3 } nvmlValueType_t;

typedef union nvmlValue_st { double dVal; int siVal; unsigned int uiVal; unsigned long ulVal; unsigned long long ullVal; signed long long sllVal; unsigned short usVal; } nvmlValue_t;
typedef struct nvmlSample_st { unsigned long long timeStamp; nvmlValue_t sampleValue; } nvmlSample_t;   // timeStamp: CPU microseconds
typedef struct nvmlMemory_st { unsigned long long total, free, used; } nvmlMemory_t;
typedef struct nvmlPciInfo_st {
  char busIdLegacy[16];
  unsigned int domain, bus, device, pciDeviceId, pciSubSystemId;
  char busId[32];                                    // "00000000:1B:00.0"
} nvmlPciInfo_t;
typedef struct nvmlEventData_st {
  nvmlDevice_t device;
  unsigned long long eventType, eventData;           // eventData = the Xid for nvmlEventTypeXidCriticalError
  unsigned int gpuInstanceId, computeInstanceId;     // 0xFFFFFFFF when the event is not tied to a MIG instance
} nvmlEventData_t;

#define nvmlEventTypeXidCriticalError 0x0000000000000008LL

nvmlReturn_t nvmlInit_v2(void);
nvmlReturn_t nvmlShutdown(void);
const char* nvmlErrorString(nvmlReturn_t result);
nvmlReturn_t nvmlSystemGetDriverVersion(char* version, unsigned int length);
nvmlReturn_t nvmlDeviceGetCount_v2(unsigned int* deviceCount);
nvmlReturn_t nvmlDeviceGetHandleByIndex_v2(unsigned int index, nvmlDevice_t* device);
nvmlReturn_t nvmlDeviceGetHandleByUUID(const char* uuid, nvmlDevice_t* device);
nvmlReturn_t nvmlDeviceGetUUID(nvmlDevice_t device, char* uuid, unsigned int length);
nvmlReturn_t nvmlDeviceGetName(nvmlDevice_t device, char* name, unsigned int length);
nvmlReturn_t nvmlDeviceGetMinorNumber(nvmlDevice_t device, unsigned int* minorNumber);
nvmlReturn_t nvmlDeviceGetPciInfo_v3(nvmlDevice_t device, nvmlPciInfo_t* pci);
nvmlReturn_t nvmlDeviceGetMemoryInfo(nvmlDevice_t device, nvmlMemory_t* memory);
nvmlReturn_t nvmlDeviceGetMigMode(nvmlDevice_t device, unsigned int* currentMode, unsigned int* pendingMode);
nvmlReturn_t nvmlDeviceGetSamples(nvmlDevice_t device, nvmlSamplingType_t type, unsigned long long lastSeenTimeStamp, nvmlValueType_t* sampleValType,
                                  unsigned int* sampleCount, nvmlSample_t* samples);
nvmlReturn_t nvmlDeviceGetSupportedEventTypes(nvmlDevice_t device, unsigned long long* eventTypes);
nvmlReturn_t nvmlDeviceRegisterEvents(nvmlDevice_t device, unsigned long long eventTypes, nvmlEventSet_t set);
nvmlReturn_t nvmlEventSetCreate(nvmlEventSet_t* set);
nvmlReturn_t nvmlEventSetWait_v2(nvmlEventSet_t set, nvmlEventData_t* data, unsigned int timeoutms);
nvmlReturn_t nvmlEventSetFree(nvmlEventSet_t set);

#ifdef __cplusplus
}
static_assert(sizeof(nvmlSample_t) == 16, "nvmlSample_t layout");
static_assert(sizeof(nvmlMemory_t) == 24, "nvmlMemory_t layout");
static_assert(sizeof(nvmlPciInfo_t) == 68, "nvmlPciInfo_t layout");
static_assert(sizeof(nvmlEventData_t) == 32, "nvmlEventData_t layout");
#endif

#endif  // __nvml_nvml_h__
