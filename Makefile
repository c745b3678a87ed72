# Top-level targets (role of reference Makefile:14-106: test / vet / presubmit / build / container*).
PY ?= python
all: build
build:                       ## compile every native artefact for sm_100a / x86-64, in-tree
	$(PY) -c "import __graft_entry__ as g; g.build()"
manifests:                   ## regenerate deploy/**.yaml from deploy/generate.py
	$(PY) deploy/generate.py
test: build                  ## CPU test-suite (GPU tests: `make test-gpu` on a B200 box)
	$(PY) -m pytest tests -x -q -m "not gpu"
conformance: build           ## the fifteen kubelet-side scenarios against both plugin implementations, as processes
	$(PY) conformance/run.py --impl native
	$(PY) conformance/run.py --impl python
test-race: build             ## the native daemons under ThreadSanitizer, then AddressSanitizer+UBSan (role of `go test -race`, reference Makefile:21)
	B200_NATIVE_SAN=thread $(PY) -m pytest tests/test_native_device_plugin.py tests/test_nri_injector.py tests/test_native_tools.py -x -q
	B200_NATIVE_SAN=address $(PY) -m pytest tests/test_native_device_plugin.py tests/test_nri_injector.py tests/test_native_tools.py -x -q
	$(MAKE) -C coll emu-tsan    # the send/recv kernel source, every CTA a host thread, under ThreadSanitizer
coverage-native: build       ## line coverage of the C++ daemons under the conformance suites (gcov)
	rm -rf build/agent-cov
	B200_NATIVE_SAN=cov $(PY) -m pytest tests/test_native_device_plugin.py tests/test_nri_injector.py tests/test_native_tools.py -q
	cd agent/native && for g in ../../build/agent-cov/*.gcda; do gcov -o ../../build/agent-cov $$g 2>/dev/null | grep -A1 "File '\(dp/\|[a-z_0-9]*\.cc\)" | grep -v "^--"; done; rm -f *.gcov
test-gpu: build
	$(PY) -m pytest tests -x -q -m gpu
presubmit: vet                ## header + style + manifest freshness
	$(PY) build_tools/boilerplate.py
	bash build_tools/check_style.sh
	$(PY) deploy/generate.py --check
	$(PY) build_tools/check_dockerfiles.py
bench:
	$(PY) bench.py --table
vet:                         ## static checks (role of reference Makefile:27-29 `go vet`): warnings-as-errors syntax pass + byte-compile
	for f in agent/native/*.cc agent/native/dp/*.cc; do $(CXX) -std=c++17 -Wall -Wextra -Werror -fsyntax-only -Iagent/native/dp $$f || exit 1; done
	$(PY) -m compileall -q container_engine_accelerators_b200 deploy bench build_tools bench.py __graft_entry__.py

# ---- images (role of reference Makefile:49-99: container, container-multi-arch, push*, partition-gpu*, nri-device-injector*,
# nvidia_persistenced_installer*, fastsocket_installer). One Dockerfile per image under docker/; IMAGE-<name> builds one,
# `containers` builds all, `push` pushes all, `containers-multi-arch` builds amd64+arm64 (GB200/GB300 nodes) with buildx.
REGISTRY ?= ghcr.io/b200-node-accelerators
TAG ?= $(shell cat VERSION 2>/dev/null || echo dev)
IMAGES := $(patsubst docker/%.Dockerfile,%,$(wildcard docker/*.Dockerfile))
MULTI_ARCH_IMAGES := device-plugin-native nri-device-injector partition-gpu persistenced topology-scheduler device-plugin
# image names as the manifests reference them (deploy/generate.py IMG): b200-<dockerfile stem> unless listed here
image_name = $(or $(IMAGE_NAME_$(1)),b200-$(1))
IMAGE_NAME_b200coll-installer := b200coll-installer
IMAGE_NAME_fastsocket-installer := fastsocket-installer
IMAGE_NAME_driver-installer-ubuntu := b200-ubuntu-driver-installer
$(addprefix image-,$(IMAGES)): image-%:
	docker build -f docker/$*.Dockerfile -t $(REGISTRY)/$(call image_name,$*):$(TAG) .
image-driver-installer-minikube:     ## same Dockerfile as the Ubuntu installer, other entrypoint
	docker build -f docker/driver-installer-ubuntu.Dockerfile --build-arg ENTRY=minikube -t $(REGISTRY)/b200-minikube-driver-installer:$(TAG) .
containers: $(addprefix image-,$(IMAGES)) image-driver-installer-minikube   ## build every image (needs docker + network)
push: containers
	for n in $(foreach i,$(IMAGES),$(call image_name,$(i))) b200-minikube-driver-installer; do docker push $(REGISTRY)/$$n:$(TAG) || exit 1; done
containers-multi-arch:       ## CPU-only images for linux/amd64 + linux/arm64 (the CUDA images are built per arch by their base image)
	for n in $(MULTI_ARCH_IMAGES); do docker buildx build --platform linux/amd64,linux/arm64 -f docker/$$n.Dockerfile -t $(REGISTRY)/b200-$$n:$(TAG) --push . || exit 1; done
device-plugin: image-device-plugin image-device-plugin-native
partition-gpu: image-partition-gpu
nri-device-injector: image-nri-device-injector
nvidia-persistenced-installer: image-persistenced
transport-installer: image-b200coll-installer image-fastsocket-installer     ## the fastsocket_installer target's role
sass:                        ## SASS listing of the collective kernels -> profiles/
	cuobjdump -sass coll/lib/libb200coll.so > profiles/libb200coll.sass
clean:
	$(MAKE) -C coll clean; $(MAKE) -C tools clean; $(MAKE) -C agent/native clean
.PHONY: all build manifests test conformance test-race coverage-native test-gpu vet presubmit bench containers push containers-multi-arch device-plugin partition-gpu nri-device-injector nvidia-persistenced-installer transport-installer sass clean
