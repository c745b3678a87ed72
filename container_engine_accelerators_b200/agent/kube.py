"""Minimal Kubernetes REST client (requests) — the image has no `kubernetes` package and the agent needs a
dozen calls: node get / status update / server-side-apply annotations / label merge-patch, event create,
pod + node list, pod replace. In-cluster config mirrors client-go's rest.InClusterConfig
(reference: pkg/gpu/nvidia/util/util.go:55-70).
"""
from __future__ import annotations

import datetime
import json
import os
from typing import Optional

import requests

SA_DIR = "/var/run/secrets/kubernetes.io/serviceaccount"


class KubeError(RuntimeError):
    def __init__(self, status: int, body: str, what: str):
        super().__init__(f"{what}: HTTP {status}: {body[:300]}")
        self.status = status


def now_rfc3339() -> str:
    return datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ")


class KubeClient:
    def __init__(self, base_url: str, token: Optional[str] = None, ca_cert: Optional[str] = None, timeout: float = 30.0):
        self.base = base_url.rstrip("/")
        self.timeout = timeout
        self.session = requests.Session()
        if token:
            self.session.headers["Authorization"] = f"Bearer {token}"
        # passed on every request: a session-level `verify` loses to REQUESTS_CA_BUNDLE / CURL_CA_BUNDLE from the environment
        # (requests merges the environment in at request level), and the cluster CA must win over the image's CA bundle
        self.verify = ca_cert if ca_cert else True
        if base_url.startswith("http://"):
            self.verify = False

    @classmethod
    def in_cluster(cls) -> "KubeClient":
        host, port = os.environ.get("KUBERNETES_SERVICE_HOST"), os.environ.get("KUBERNETES_SERVICE_PORT", "443")
        if not host:
            raise KubeError(0, "", "not running in a cluster (KUBERNETES_SERVICE_HOST unset)")
        with open(os.path.join(SA_DIR, "token")) as f:
            token = f.read().strip()
        if ":" in host:
            host = f"[{host}]"
        return cls(f"https://{host}:{port}", token, os.path.join(SA_DIR, "ca.crt"))

    @classmethod
    def from_env(cls, url: str = "") -> "KubeClient":
        """Where every daemon gets its client: an explicit URL, else B200_KUBE_URL (+ B200_KUBE_TOKEN_FILE / B200_KUBE_CA_FILE: the same
        overrides the native binary honours, agent/native/dp/kube.hpp), else the pod's in-cluster configuration."""
        url = url or os.environ.get("B200_KUBE_URL", "")
        if not url:
            return cls.in_cluster()
        token = None
        token_file = os.environ.get("B200_KUBE_TOKEN_FILE", "")
        if token_file:
            with open(token_file) as f:
                token = f.read().strip()
        return cls(url, token, os.environ.get("B200_KUBE_CA_FILE") or None)

    # ------------------------------------------------------------------ plumbing
    def _req(self, method: str, path: str, what: str, body=None, headers=None, params=None, raw_body: Optional[str] = None):
        h = dict(headers or {})
        data = raw_body
        if body is not None:
            data = json.dumps(body)
            h.setdefault("Content-Type", "application/json")
        r = self.session.request(method, self.base + path, data=data, headers=h, params=params, timeout=self.timeout, verify=self.verify)
        if r.status_code >= 300:
            raise KubeError(r.status_code, r.text, what)
        return r.json() if r.content else {}

    # ------------------------------------------------------------------ nodes
    def get_node(self, name: str) -> dict:
        return self._req("GET", f"/api/v1/nodes/{name}", f"get node {name}")

    def list_nodes(self) -> list:
        return self._req("GET", "/api/v1/nodes", "list nodes").get("items", [])

    def update_node_status(self, node: dict) -> dict:
        name = node["metadata"]["name"]
        return self._req("PUT", f"/api/v1/nodes/{name}/status", f"update node {name} status", body=node)

    def apply_node_annotations(self, name: str, annotations: dict, field_manager: str, force: bool = True) -> dict:
        """Server-side apply: owns only the listed annotations, leaves every other one alone."""
        patch = {"apiVersion": "v1", "kind": "Node", "metadata": {"name": name, "annotations": annotations}}
        return self._req("PATCH", f"/api/v1/nodes/{name}", f"apply annotations on node {name}", raw_body=json.dumps(patch),
                         headers={"Content-Type": "application/apply-patch+yaml"}, params={"fieldManager": field_manager, "force": "true" if force else "false"})

    def patch_node_labels(self, name: str, labels: dict) -> dict:
        return self._req("PATCH", f"/api/v1/nodes/{name}", f"label node {name}", raw_body=json.dumps({"metadata": {"labels": labels}}),
                         headers={"Content-Type": "application/merge-patch+json"})

    # ------------------------------------------------------------------ events
    def create_event(self, namespace: str, involved: dict, event_type: str, reason: str, message: str, component: str) -> dict:
        ts = now_rfc3339()
        body = {"apiVersion": "v1", "kind": "Event", "metadata": {"generateName": f"{involved.get('name', 'obj')}.", "namespace": namespace},
                "involvedObject": involved, "type": event_type, "reason": reason, "message": message, "source": {"component": component},
                "firstTimestamp": ts, "lastTimestamp": ts, "count": 1}
        return self._req("POST", f"/api/v1/namespaces/{namespace}/events", "create event", body=body)

    # ------------------------------------------------------------------ pods
    def list_pods(self, field_selector: str = "", namespace: str = "") -> list:
        path = f"/api/v1/namespaces/{namespace}/pods" if namespace else "/api/v1/pods"
        return self._req("GET", path, "list pods", params={"fieldSelector": field_selector} if field_selector else None).get("items", [])

    def get_pod(self, namespace: str, name: str) -> dict:
        return self._req("GET", f"/api/v1/namespaces/{namespace}/pods/{name}", f"get pod {namespace}/{name}")

    def replace_pod(self, namespace: str, name: str, pod: dict) -> dict:
        return self._req("PUT", f"/api/v1/namespaces/{namespace}/pods/{name}", f"replace pod {namespace}/{name}", body=pod)
