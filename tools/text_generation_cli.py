#!/usr/bin/env python
"""Interactive client of the text-generation server (reference ``tools/text_generation_cli.py``): reads a prompt, PUTs it to
``http://host:port/api`` and prints the completion.

    python tools/text_generation_cli.py localhost:5000 [--tokens 64] [--temperature 1.0] [--top-k 0] [--top-p 0.0]
"""
import argparse
import json
import sys
import urllib.request


def request(url: str, prompt: str, tokens: int, **sampling) -> str:
    body = {"prompts": [prompt], "tokens_to_generate": tokens, **{k: v for k, v in sampling.items() if v is not None}}
    req = urllib.request.Request(url, data=json.dumps(body).encode(), headers={"Content-Type": "application/json"}, method="PUT")
    with urllib.request.urlopen(req, timeout=600) as r:
        out = json.loads(r.read().decode())
    if "text" not in out:
        raise RuntimeError(out.get("message", str(out)))
    return out["text"][0]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("server", help="host:port")
    ap.add_argument("--tokens", type=int, default=64)
    ap.add_argument("--temperature", type=float, default=None)
    ap.add_argument("--top-k", type=int, default=None)
    ap.add_argument("--top-p", type=float, default=None)
    ap.add_argument("--prompt", default=None, help="one-shot instead of the interactive loop")
    a = ap.parse_args(argv)
    url = f"http://{a.server}/api"
    kw = dict(temperature=a.temperature, top_k=a.top_k, top_p=a.top_p)
    if a.prompt is not None:
        print(request(url, a.prompt, a.tokens, **kw))
        return
    while True:
        try:
            prompt = input("Enter prompt: ")
        except EOFError:
            break
        print("Megatron Response:\n" + request(url, prompt, a.tokens, **kw), file=sys.stdout)


if __name__ == "__main__":
    main()
