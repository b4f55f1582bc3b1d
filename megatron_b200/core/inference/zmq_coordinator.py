"""ZMQ serving plane: clients ⇄ coordinator ⇄ one dynamic engine per data-parallel replica (reference ``inference/data_parallel_inference_coordinator.py``,
``inference/inference_client.py``, ``inference/headers.py``, and the engine-side loop of ``engines/dynamic_engine.py`` ``run_engine_with_coordinator``).

One ROUTER socket in the coordinator; engines and clients are DEALERs that introduce themselves (``ENGINE_CONNECT`` / ``CLIENT_CONNECT``).  Messages are
msgpack lists ``[header, ...]``:

* client → coordinator ``SUBMIT [client_req_id, prompt_tokens, sampling]`` → routed to the unpaused engine with the least outstanding tokens (prompt + budgeted
  generation) as ``SUBMIT [global_id, prompt_tokens, sampling]``;
* engine → coordinator ``REPLY [global_id, generated_tokens, log_probs, ttft]`` → forwarded to the client that owns the request as ``REPLY [client_req_id, …]``;
* control: ``PAUSE`` / ``UNPAUSE`` (engines finish the step they are in and stop stepping — weights can be swapped, e.g. RL refit), ``STOP`` (engines drain
  and exit, then the coordinator exits), ``STATS`` (per-engine outstanding tokens / served counts).

The engine loop polls its socket without blocking while it has work, so new requests join the continuous batch at the next step."""
from __future__ import annotations

import enum
import itertools
import threading
import time
from typing import Dict, List, Optional, Tuple

import msgpack
import zmq

from .sampling import SamplingParams


class Headers(enum.IntEnum):
    ENGINE_CONNECT = 1
    CLIENT_CONNECT = 2
    ACK = 3
    SUBMIT = 4
    REPLY = 5
    PAUSE = 6
    UNPAUSE = 7
    STOP = 8
    STATS = 9
    ENGINE_PAUSED = 10


_SP_FIELDS = ("temperature", "top_k", "top_p", "num_tokens_to_generate", "return_log_probs", "stop_token_ids", "seed")


def _sp_to_dict(sp: SamplingParams) -> dict:
    return {k: (list(getattr(sp, k)) if k == "stop_token_ids" else getattr(sp, k)) for k in _SP_FIELDS if hasattr(sp, k)}


def _sp_from_dict(d: dict) -> SamplingParams:
    sp = SamplingParams()
    for k, v in d.items():
        if hasattr(sp, k):
            setattr(sp, k, type(getattr(sp, k))(v) if k == "stop_token_ids" and getattr(sp, k) is not None else v)
    return sp


class ZMQCoordinator:
    def __init__(self, port: int, num_engines: int, host: str = "127.0.0.1"):
        self.addr, self.num_engines = f"tcp://{host}:{port}", num_engines
        self.ctx = zmq.Context.instance()
        self.sock = self.ctx.socket(zmq.ROUTER)
        self.sock.bind(self.addr)
        self.engines: List[bytes] = []
        self.load: Dict[bytes, int] = {}
        self.served: Dict[bytes, int] = {}
        self.paused = False
        self.paused_acks = 0
        self._gid = itertools.count()
        self.owner: Dict[int, Tuple[bytes, int, bytes, int]] = {}     # global id -> (client identity, client request id, engine identity, cost)
        self.pending: List[Tuple[bytes, list]] = []                     # submissions that arrived before every engine had connected

    def _send(self, ident: bytes, msg: list) -> None:
        self.sock.send_multipart([ident, msgpack.packb(msg, use_bin_type=True)])

    def _route(self, client: bytes, body: list) -> None:
        creq, tokens, sampling = body
        eng = min(self.engines, key=lambda e: (self.load[e], self.engines.index(e)))
        gid = next(self._gid)
        cost = len(tokens) + int(sampling.get("num_tokens_to_generate", 0))
        self.owner[gid] = (client, creq, eng, cost)
        self.load[eng] += cost
        self._send(eng, [int(Headers.SUBMIT), gid, tokens, sampling])

    def run(self) -> None:
        """Serve until STOP.  Engines must connect first (``num_engines`` of them); client traffic that arrives earlier is queued."""
        stopping_client: Optional[bytes] = None
        while True:
            ident, raw = self.sock.recv_multipart()
            msg = msgpack.unpackb(raw, raw=False)
            h = Headers(msg[0])
            if h == Headers.ENGINE_CONNECT:
                self.engines.append(ident)
                self.load[ident], self.served[ident] = 0, 0
                self._send(ident, [int(Headers.ACK)])
                if len(self.engines) == self.num_engines:
                    for c, body in self.pending:
                        self._route(c, body)
                    self.pending.clear()
            elif h == Headers.CLIENT_CONNECT:
                self._send(ident, [int(Headers.ACK), len(self.engines)])
            elif h == Headers.SUBMIT:
                if len(self.engines) < self.num_engines:
                    self.pending.append((ident, msg[1:]))
                else:
                    self._route(ident, msg[1:])
            elif h == Headers.REPLY:
                gid = msg[1]
                client, creq, eng, cost = self.owner.pop(gid)
                self.load[eng] -= cost
                self.served[eng] += 1
                self._send(client, [int(Headers.REPLY), creq] + msg[2:])
            elif h in (Headers.PAUSE, Headers.UNPAUSE):
                self.paused = h == Headers.PAUSE
                self.paused_acks = 0
                for e in self.engines:
                    self._send(e, [int(h)])
                self._send(ident, [int(Headers.ACK)])
            elif h == Headers.ENGINE_PAUSED:
                self.paused_acks += 1
            elif h == Headers.STATS:
                self._send(ident, [int(Headers.STATS), [self.load[e] for e in self.engines], [self.served[e] for e in self.engines], self.paused, self.paused_acks])
            elif h == Headers.STOP:
                for e in self.engines:
                    self._send(e, [int(Headers.STOP)])
                self._send(ident, [int(Headers.ACK)])
                stopping_client = ident
                break
        self.sock.close(linger=200)
        del stopping_client


class EngineWorker:
    """Drives one ``DynamicInferenceEngine`` from the coordinator's socket (the reference's ``run_engine_with_coordinator`` loop)."""

    def __init__(self, engine, port: int, host: str = "127.0.0.1", idle_poll_ms: int = 20):
        self.engine, self.addr, self.idle_poll_ms = engine, f"tcp://{host}:{port}", idle_poll_ms
        self.local_to_global: Dict[int, int] = {}
        self.paused = False
        self.steps_while_paused = 0

    def run(self) -> None:
        ctx = zmq.Context.instance()
        sock = ctx.socket(zmq.DEALER)
        sock.connect(self.addr)
        sock.send(msgpack.packb([int(Headers.ENGINE_CONNECT)]))
        assert msgpack.unpackb(sock.recv(), raw=False)[0] == Headers.ACK
        stop = False
        while True:
            busy = self.engine.has_unfinished() and not self.paused
            while sock.poll(0 if busy else self.idle_poll_ms):
                msg = msgpack.unpackb(sock.recv(), raw=False)
                h = Headers(msg[0])
                if h == Headers.SUBMIT:
                    gid, tokens, sampling = msg[1:]
                    self.local_to_global[self.engine.add_request(list(tokens), _sp_from_dict(sampling))] = gid
                elif h == Headers.PAUSE:
                    self.paused = True
                    sock.send(msgpack.packb([int(Headers.ENGINE_PAUSED)]))
                elif h == Headers.UNPAUSE:
                    self.paused = False
                elif h == Headers.STOP:
                    stop = True
                busy = self.engine.has_unfinished() and not self.paused
            if stop and not self.engine.has_unfinished():
                break
            if self.engine.has_unfinished() and (not self.paused or stop):
                for req in self.engine.step():
                    sock.send(msgpack.packb([int(Headers.REPLY), self.local_to_global.pop(req.request_id), list(req.generated_tokens), list(req.log_probs), req.ttft]))
        sock.close(linger=200)


class ZMQInferenceClient:
    """Submit / collect against the coordinator; ``generate`` is the blocking convenience wrapper."""

    def __init__(self, port: int, host: str = "127.0.0.1"):
        self.sock = zmq.Context.instance().socket(zmq.DEALER)
        self.sock.connect(f"tcp://{host}:{port}")
        self._ids = itertools.count()
        self.results: Dict[int, dict] = {}
        self._rpc([int(Headers.CLIENT_CONNECT)])

    def _rpc(self, msg: list, expect: Headers = Headers.ACK, timeout_ms: int = 30000) -> list:
        self.sock.send(msgpack.packb(msg, use_bin_type=True))
        deadline = time.time() + timeout_ms / 1e3
        while time.time() < deadline:
            if self.sock.poll(50):
                m = msgpack.unpackb(self.sock.recv(), raw=False)
                if Headers(m[0]) == expect:
                    return m
                self._store(m)
        raise TimeoutError(f"no {expect.name} from the coordinator")

    def _store(self, m: list) -> None:
        if Headers(m[0]) == Headers.REPLY:
            self.results[m[1]] = {"generated_tokens": m[2], "log_probs": m[3], "ttft": m[4]}

    def submit(self, prompt_tokens: List[int], sampling_params: Optional[SamplingParams] = None) -> int:
        rid = next(self._ids)
        self.sock.send(msgpack.packb([int(Headers.SUBMIT), rid, list(prompt_tokens), _sp_to_dict(sampling_params or SamplingParams())], use_bin_type=True))
        return rid

    def collect(self, ids: List[int], timeout_s: float = 120.0) -> Dict[int, dict]:
        deadline = time.time() + timeout_s
        while any(i not in self.results for i in ids):
            if time.time() > deadline:
                raise TimeoutError(f"requests {[i for i in ids if i not in self.results]} did not finish")
            if self.sock.poll(50):
                self._store(msgpack.unpackb(self.sock.recv(), raw=False))
        return {i: self.results[i] for i in ids}

    def generate(self, prompts: List[List[int]], sampling_params: Optional[SamplingParams] = None, timeout_s: float = 120.0) -> List[List[int]]:
        ids = [self.submit(p, sampling_params) for p in prompts]
        out = self.collect(ids, timeout_s)
        return [out[i]["generated_tokens"] for i in ids]

    def pause_engines(self) -> None:
        self._rpc([int(Headers.PAUSE)])

    def unpause_engines(self) -> None:
        self._rpc([int(Headers.UNPAUSE)])

    def stats(self) -> dict:
        m = self._rpc([int(Headers.STATS)], expect=Headers.STATS)
        return {"outstanding_tokens": m[1], "served": m[2], "paused": m[3], "paused_acks": m[4]}

    def stop(self) -> None:
        self._rpc([int(Headers.STOP)])
        self.sock.close(linger=200)


def start_in_threads(engines: List[object], port: int) -> Tuple[threading.Thread, List[threading.Thread]]:
    """Coordinator + one worker thread per engine in THIS process (tests, single-process serving); multi-process deployments run ``ZMQCoordinator.run`` in the
    launcher and ``EngineWorker.run`` on rank 0 of every model-parallel replica."""
    coord = ZMQCoordinator(port, len(engines))
    ct = threading.Thread(target=coord.run, daemon=True)
    ct.start()
    wts = [threading.Thread(target=EngineWorker(e, port).run, daemon=True) for e in engines]
    for t in wts:
        t.start()
    return ct, wts
