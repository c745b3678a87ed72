# Top-level targets (role of reference Makefile:14-106: test / vet / presubmit / build / container*).
PY ?= python
all: build
build:                       ## compile every native artefact for sm_100a / x86-64, in-tree
	$(PY) -c "import __graft_entry__ as g; g.build()"
manifests:                   ## regenerate deploy/**.yaml from deploy/generate.py
	$(PY) deploy/generate.py
test: build                  ## CPU test-suite (GPU tests: `make test-gpu` on a B200 box)
	$(PY) -m pytest tests -x -q -m "not gpu"
test-gpu: build
	$(PY) -m pytest tests -x -q -m gpu
presubmit:                   ## header + style + manifest freshness
	$(PY) build_tools/boilerplate.py
	bash build_tools/check_style.sh
	$(PY) deploy/generate.py --check
bench:
	$(PY) bench.py --table
containers:                  ## build every image under docker/ (needs docker + network)
	for f in docker/*.Dockerfile; do n=$$(basename $$f .Dockerfile); docker build -f $$f -t b200-$$n:dev . || exit 1; done
sass:                        ## SASS listing of the collective kernels -> profiles/
	cuobjdump -sass coll/lib/libb200coll.so > profiles/libb200coll.sass
clean:
	$(MAKE) -C coll clean; $(MAKE) -C tools clean; $(MAKE) -C agent/native clean
.PHONY: all build manifests test test-gpu presubmit bench containers sass clean
