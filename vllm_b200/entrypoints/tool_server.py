"""Server-side tools for the Responses API: an MCP client and the model <-> tool loop.

Behavioural reference: vllm/entrypoints/mcp/tool_server.py (``MCPToolServer``: connect to MCP servers, list their
tools, open a session per call), entrypoints/openai/responses/serving.py (the loop that feeds tool results back to the
model until it answers in text) and the ``--tool-server`` flag.

The MCP "streamable HTTP" transport is plain JSON-RPC 2.0 over POST: the reply is either ``application/json`` or a
short ``text/event-stream`` whose ``data:`` lines carry the JSON-RPC message; a ``Mcp-Session-Id`` header returned by
``initialize`` is echoed on later calls.  That is all a client needs, so it is implemented here on aiohttp rather than
depending on the ``mcp`` package.
"""
from __future__ import annotations

import itertools
import json
from typing import Any, Awaitable, Callable, Optional

PROTOCOL_VERSION = "2025-03-26"


class MCPError(RuntimeError):
    pass


class MCPClient:
    def __init__(self, url: str, label: Optional[str] = None, headers: Optional[dict[str, str]] = None,
                 timeout: float = 30.0) -> None:
        self.url, self.label = url, label or url
        self.headers = dict(headers or {})
        self.timeout = timeout
        self.session_id: Optional[str] = None
        self.server_info: dict[str, Any] = {}
        self._ids = itertools.count(1)
        self._tools: Optional[list[dict[str, Any]]] = None

    async def _post(self, payload: dict[str, Any]) -> Optional[dict[str, Any]]:
        import aiohttp
        h = {"Content-Type": "application/json", "Accept": "application/json, text/event-stream", **self.headers}
        if self.session_id:
            h["Mcp-Session-Id"] = self.session_id
        async with aiohttp.ClientSession(timeout=aiohttp.ClientTimeout(total=self.timeout)) as s:
            async with s.post(self.url, json=payload, headers=h) as r:
                if r.status >= 400:
                    raise MCPError(f"MCP server {self.label}: HTTP {r.status}")
                sid = r.headers.get("Mcp-Session-Id")
                if sid:
                    self.session_id = sid
                body = await r.text()
                if "id" not in payload:                       # notification: no reply expected
                    return None
                if "text/event-stream" in r.headers.get("Content-Type", ""):
                    msg = None
                    for line in body.splitlines():
                        if line.startswith("data:"):
                            d = json.loads(line[5:].strip())
                            if d.get("id") == payload["id"]:
                                msg = d
                    if msg is None:
                        raise MCPError(f"MCP server {self.label}: no reply to request {payload['id']} in event stream")
                else:
                    msg = json.loads(body)
        if "error" in msg:
            raise MCPError(f"MCP server {self.label}: {msg['error'].get('message', msg['error'])}")
        return msg.get("result", {})

    async def request(self, method: str, params: Optional[dict[str, Any]] = None) -> dict[str, Any]:
        return await self._post({"jsonrpc": "2.0", "id": next(self._ids), "method": method, "params": params or {}})

    async def initialize(self) -> None:
        if self.server_info:
            return
        res = await self.request("initialize", {"protocolVersion": PROTOCOL_VERSION, "capabilities": {},
                                                "clientInfo": {"name": "vllm_b200", "version": "0.1"}})
        self.server_info = res.get("serverInfo", {}) or {"name": self.label}
        await self._post({"jsonrpc": "2.0", "method": "notifications/initialized"})

    async def list_tools(self) -> list[dict[str, Any]]:
        if self._tools is None:
            await self.initialize()
            tools, cursor = [], None
            while True:
                res = await self.request("tools/list", {"cursor": cursor} if cursor else {})
                tools += res.get("tools", [])
                cursor = res.get("nextCursor")
                if not cursor:
                    break
            self._tools = tools
        return self._tools

    async def call_tool(self, name: str, arguments: dict[str, Any]) -> str:
        await self.initialize()
        res = await self.request("tools/call", {"name": name, "arguments": arguments})
        text = "".join(c.get("text", "") for c in res.get("content", []) if c.get("type") == "text")
        if res.get("isError"):
            raise MCPError(text or f"tool {name} failed")
        if not text and res.get("structuredContent") is not None:
            text = json.dumps(res["structuredContent"])
        return text


class ToolServer:
    """The set of MCP servers one API server can reach (``--tool-server url[,url...]`` plus the per-request
    ``{"type": "mcp", "server_url": ...}`` tools)."""

    def __init__(self, urls: Optional[list[str]] = None) -> None:
        self.clients: dict[str, MCPClient] = {}
        for u in urls or []:
            self.add(u if "://" in u else f"http://{u}")

    def add(self, url: str, label: Optional[str] = None, headers: Optional[dict[str, str]] = None) -> MCPClient:
        key = label or url
        c = self.clients.get(key)
        if c is None or c.url != url:
            c = self.clients[key] = MCPClient(url, label or url, headers)
        return c

    async def resolve(self, tools: Optional[list[dict[str, Any]]]) -> tuple[list[dict[str, Any]], dict[str, tuple[MCPClient, str]],
                                                                           list[dict[str, Any]]]:
        """Request ``tools`` → (chat-completions function tools, {function name: (client, remote tool name)},
        ``mcp_list_tools`` output items).  ``function`` tools pass through; ``mcp`` tools are expanded to the functions
        their server advertises (filtered by ``allowed_tools``)."""
        fns: list[dict[str, Any]] = []
        route: dict[str, tuple[MCPClient, str]] = {}
        items: list[dict[str, Any]] = []
        wanted = list(tools or [])
        # servers configured on the command line are available to every request
        have = {t.get("server_url") for t in wanted if t.get("type") == "mcp"}
        for key, c in self.clients.items():
            if c.url not in have and c.label == c.url:
                wanted.append({"type": "mcp", "server_url": c.url, "server_label": key})
        for t in wanted:
            typ = t.get("type")
            if typ == "function":
                f = t.get("function") or {k: t[k] for k in ("name", "description", "parameters", "strict") if k in t}
                fns.append({"type": "function", "function": f})
            elif typ == "mcp":
                c = self.add(t["server_url"], t.get("server_label"), t.get("headers"))
                allowed = t.get("allowed_tools")
                if isinstance(allowed, dict):
                    allowed = allowed.get("tool_names")
                listed = []
                for mt in await c.list_tools():
                    if allowed and mt["name"] not in allowed:
                        continue
                    fname = mt["name"] if mt["name"] not in route else f"{c.label}__{mt['name']}"
                    route[fname] = (c, mt["name"])
                    fns.append({"type": "function", "function": {
                        "name": fname, "description": mt.get("description", ""),
                        "parameters": mt.get("inputSchema") or {"type": "object", "properties": {}}}})
                    listed.append({"name": mt["name"], "description": mt.get("description"),
                                   "input_schema": mt.get("inputSchema")})
                items.append({"type": "mcp_list_tools", "server_label": c.label, "tools": listed})
            # other built-in tool types (web_search_preview, code_interpreter, ...) need an MCP server that offers them
        return fns, route, items


ChatFn = Callable[[list[dict[str, Any]], list[dict[str, Any]]], Awaitable[dict[str, Any]]]


async def run_tool_loop(chat: ChatFn, messages: list[dict[str, Any]], fns: list[dict[str, Any]],
                        route: dict[str, tuple[MCPClient, str]], max_turns: int = 8
                        ) -> tuple[dict[str, Any], list[dict[str, Any]], dict[str, int]]:
    """Calls ``chat(messages, tools)`` (→ chat-completions assistant ``message`` + ``usage``) until the model stops
    asking for server-side tools.  Returns (last message, output items produced on the way, summed usage).
    Tool calls that are *not* routed to an MCP server (plain ``function`` tools) end the loop: the client runs those."""
    items: list[dict[str, Any]] = []
    usage = {"prompt_tokens": 0, "completion_tokens": 0, "total_tokens": 0}
    msg: dict[str, Any] = {}
    for _ in range(max_turns):
        res = await chat(messages, fns)
        for k in usage:
            usage[k] += int((res.get("usage") or {}).get(k, 0))
        msg = res["message"]
        calls = msg.get("tool_calls") or []
        if not calls or any(c["function"]["name"] not in route for c in calls):
            break
        messages.append({"role": "assistant", "content": msg.get("content") or "", "tool_calls": calls})
        for c in calls:
            client, remote = route[c["function"]["name"]]
            raw = c["function"].get("arguments") or "{}"
            item = {"type": "mcp_call", "id": f"mcp_{c['id']}", "server_label": client.label, "name": remote,
                    "arguments": raw, "output": None, "error": None}
            try:
                out = await client.call_tool(remote, json.loads(raw) if isinstance(raw, str) else raw)
                item["output"] = out
            except (MCPError, ValueError, OSError) as e:
                out = f"error: {e}"
                item["error"] = str(e)
            items.append(item)
            messages.append({"role": "tool", "tool_call_id": c["id"], "content": out})
    return msg, items, usage


def input_items_to_messages(items: list[dict[str, Any]]) -> list[dict[str, Any]]:
    """Responses-API input items → chat messages, including the function-call round trip
    (``function_call`` → assistant tool call, ``function_call_output`` → tool message)."""
    msgs: list[dict[str, Any]] = []
    for it in items:
        typ = it.get("type", "message")
        if typ == "function_call":
            call = {"id": it.get("call_id") or it.get("id"), "type": "function",
                    "function": {"name": it["name"], "arguments": it.get("arguments") or "{}"}}
            if msgs and msgs[-1]["role"] == "assistant" and msgs[-1].get("tool_calls") is not None:
                msgs[-1]["tool_calls"].append(call)
            else:
                msgs.append({"role": "assistant", "content": "", "tool_calls": [call]})
        elif typ == "function_call_output":
            out = it.get("output")
            msgs.append({"role": "tool", "tool_call_id": it.get("call_id"),
                         "content": out if isinstance(out, str) else json.dumps(out)})
        elif typ in ("message", None) or "role" in it:
            c = it.get("content")
            if isinstance(c, list):
                c = "".join(p.get("text", "") for p in c if isinstance(p, dict))
            msgs.append({"role": it.get("role", "user"), "content": c})
        # reasoning / mcp_call / mcp_list_tools items of an earlier turn carry nothing the template needs
    return msgs
