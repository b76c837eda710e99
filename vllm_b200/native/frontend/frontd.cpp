// vllm_b200 native front-end daemon: HTTP (epoll) <-> engine-core processes (ZeroMQ + msgpack), no Python in the
// request path.  Role equivalent to the reference's Rust front-end (rust/src: engine-core-client + llm token-in/out
// facade + incremental detokeniser + axum server), redesigned as one small C++ binary:
//
//   HTTP thread   : epoll loop; parses requests, builds EngineCoreRequest msgpack, sends on the ROUTER socket,
//                   owns every client socket (reads, writes, disconnect → abort).
//   output thread : blocks on the PULL socket, decodes EngineCoreOutputs, detokenises incrementally (id → bytes
//                   table written by the Python launcher, UTF-8 boundary hold-back, stop strings), formats the
//                   JSON / SSE payloads and hands them to the HTTP thread through an eventfd.
//
// Wire protocol: vllm_b200/engine/core.py (ROUTER/DEALER with 2-byte engine identity, frames [kind, payload];
// outputs [client_index, EngineCoreOutputs]).  libzmq is dlopen'ed (pyzmq ships the shared object but no headers).
#include <arpa/inet.h>
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <signal.h>
#include <sys/epoll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "tokenizer.hpp"
#include "wire.hpp"

using wire::Val;

// ------------------------------------------------------------------------------------------------ libzmq (dlopen)
namespace z {
constexpr int PULL = 7, ROUTER = 6, SNDMORE = 2, DONTWAIT = 1, RCVTIMEO = 27, LINGER = 17, RCVMORE = 13, ROUTER_MANDATORY = 33;
void* (*ctx_new)();
int (*ctx_term)(void*);
void* (*socket)(void*, int);
int (*close)(void*);
int (*bind)(void*, const char*);
int (*setsockopt)(void*, int, const void*, size_t);
int (*getsockopt)(void*, int, void*, size_t*);
int (*send)(void*, const void*, size_t, int);
int (*msg_init)(void*);
int (*msg_recv)(void*, void*, int);
void* (*msg_data)(void*);
size_t (*msg_size)(void*);
int (*msg_close)(void*);
int (*zerrno)();
const char* (*strerror)(int);
struct Msg { alignas(8) unsigned char _[64]; };

void load(const char* path) {
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "frontd: cannot load %s: %s\n", path, dlerror()); exit(2); }
#define SYM(var, name) *(void**)(&var) = dlsym(h, name); if (!var) { fprintf(stderr, "frontd: %s missing in libzmq\n", name); exit(2); }
  SYM(ctx_new, "zmq_ctx_new") SYM(ctx_term, "zmq_ctx_term") SYM(socket, "zmq_socket") SYM(close, "zmq_close") SYM(bind, "zmq_bind")
  SYM(setsockopt, "zmq_setsockopt") SYM(getsockopt, "zmq_getsockopt") SYM(send, "zmq_send") SYM(msg_init, "zmq_msg_init")
  SYM(msg_recv, "zmq_msg_recv") SYM(msg_data, "zmq_msg_data") SYM(msg_size, "zmq_msg_size") SYM(msg_close, "zmq_msg_close")
  SYM(zerrno, "zmq_errno") SYM(strerror, "zmq_strerror")
#undef SYM
}

// one multipart message; false on timeout
bool recv_multipart(void* sock, std::vector<std::string>& frames) {
  frames.clear();
  while (true) {
    Msg m; msg_init(&m);
    if (msg_recv(&m, sock, 0) < 0) { msg_close(&m); return false; }
    frames.emplace_back((const char*)msg_data(&m), msg_size(&m));
    msg_close(&m);
    int more = 0; size_t sz = sizeof more;
    getsockopt(sock, RCVMORE, &more, &sz);
    if (!more) return true;
  }
}
}  // namespace z

// ------------------------------------------------------------------------------------------------ configuration
struct Config {
  std::string host = "127.0.0.1", in_addr, out_addr, model = "model", vocab_path, libzmq, api_key, chat_template = "plain";
  int port = 8000, num_engines = 1;
  int64_t max_model_len = 8192, eos = -1;
};

static double now_s() { return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count(); }

// ------------------------------------------------------------------------------------------------ detokeniser table
struct Vocab {
  std::vector<std::string> bytes;
  std::vector<uint8_t> special;
  bool load(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    uint32_t n = 0;
    if (fread(&n, 4, 1, f) != 1) { fclose(f); return false; }
    bytes.resize(n); special.resize(n);
    for (uint32_t i = 0; i < n; i++) {
      uint8_t fl; uint32_t len;
      if (fread(&fl, 1, 1, f) != 1 || fread(&len, 4, 1, f) != 1) { fclose(f); return false; }
      special[i] = fl; bytes[i].resize(len);
      if (len && fread(&bytes[i][0], 1, len, f) != len) { fclose(f); return false; }
    }
    char magic[4];
    if (fread(magic, 1, 4, f) == 4 && !memcmp(magic, "ENC1", 4)) {
      if (!enc.load(f)) fprintf(stderr, "frontd: encoder section of %s is damaged; text prompts disabled\n", path.c_str());
    }
    fclose(f);
    return true;
  }
  tok::Encoder enc;       // optional: present when the launcher recognised the tokenizer's BPE scheme
};

// Moves every complete character of `pending` to `text`; malformed bytes become U+FFFD (what Python's
// errors="replace" does), a trailing incomplete-but-plausible sequence stays in `pending` unless `flush`.
static void utf8_drain(std::string& pending, std::string& text, bool flush) {
  size_t i = 0, n = pending.size();
  while (i < n) {
    unsigned char c = pending[i];
    size_t need = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 0;
    if (need == 0) { text += "\xef\xbf\xbd"; i++; continue; }
    size_t have = 1;
    while (have < need && i + have < n && (static_cast<unsigned char>(pending[i + have]) & 0xc0) == 0x80) have++;
    if (have == need) { text.append(pending, i, need); i += need; continue; }
    if (i + have == n && !flush) break;                                  // may still be completed by the next token
    text += "\xef\xbf\xbd"; i += have;
  }
  pending.erase(0, i);
}

// ------------------------------------------------------------------------------------------------ state
struct Conn {
  int fd = -1;
  uint64_t id = 0;
  std::string in, out;
  bool close_after_flush = false, streaming = false;
  std::string rid;                         // request currently owning this connection ("" = idle)
};

struct Req {
  std::string rid, model, pending, text;   // pending: bytes not yet on a UTF-8 boundary
  uint64_t conn_id = 0;
  bool stream = false, token_api = false, detok = true, include_stop = false, skip_special = true, done = false, include_usage = false;
  bool chat = false, sent_role = false;     // /v1/chat/completions: chat.completion(.chunk) objects, role in the first delta
  std::vector<int64_t> out_ids;
  std::vector<std::string> stops;
  size_t n_prompt = 0, sent_text = 0, sent_ids = 0, hold = 0;
  int64_t created = 0;
  Val kv_params;
};

struct Outbox { uint64_t conn_id; std::string data; bool end_of_response; bool close_conn; };

struct Server {
  Config cfg;
  Vocab vocab;
  void *zctx = nullptr, *router = nullptr, *pull = nullptr;
  int ep = -1, lfd = -1, evfd = -1;
  std::mutex mu;                                         // guards reqs, outbox, aborts, stats
  std::unordered_map<std::string, std::unique_ptr<Req>> reqs;
  std::deque<Outbox> outbox;
  std::deque<std::string> aborts;
  std::unordered_map<uint64_t, std::unique_ptr<Conn>> conns;   // HTTP thread only
  std::unordered_map<int, uint64_t> fd2conn;
  uint64_t next_conn = 1, next_req = 1;
  int rr = 0;
  std::atomic<bool> stop{false};
  // metrics
  std::atomic<uint64_t> n_requests{0}, n_finished{0}, n_gen_tokens{0}, n_prompt_tokens{0}, n_aborted{0};
  int64_t running = 0, waiting = 0; double kv_usage = 0;

  void wake() { uint64_t one = 1; (void)!write(evfd, &one, 8); }

  // ---------------------------------------------------------------- engine I/O
  void send_engine(const char* kind, const std::string& payload) {
    std::string ident(2, '\0');
    ident[0] = char(rr % cfg.num_engines); rr++;
    z::send(router, ident.data(), 2, z::SNDMORE);
    z::send(router, kind, 1, z::SNDMORE);
    z::send(router, payload.data(), payload.size(), 0);
  }
  void send_abort(const std::string& rid) {
    wire::MpWriter w; w.array(1); w.str(rid);
    for (int e = 0; e < cfg.num_engines; e++) {
      std::string ident(2, '\0'); ident[0] = char(e);
      z::send(router, ident.data(), 2, z::SNDMORE);
      z::send(router, "B", 1, z::SNDMORE);
      z::send(router, w.o.data(), w.o.size(), 0);
    }
  }

  void handshake() {
    int ready = 0;
    std::vector<std::string> fr;
    while (ready < cfg.num_engines && !stop) {
      if (!z::recv_multipart(router, fr)) continue;
      if (fr.size() < 2 || fr[1] == "HELLO") continue;
      try {
        wire::MpReader r(fr[1].data(), fr[1].size());
        Val v = r.read();
        const Val* st = v.get("status");
        if (st && st->s == "READY") ready++;
      } catch (...) {}
    }
  }

  // ---------------------------------------------------------------- response formatting
  static const char* finish_name(int64_t f) { return f == 0 ? "stop" : f == 1 ? "length" : "abort"; }

  // chat.completion / chat.completion.chunk (reference: entrypoints/openai/chat_completion/protocol.py)
  std::string chat_json(Req& r, const std::string& text, const char* finish, bool chunk, bool usage, bool no_choices = false) {
    std::string o = "{\"id\":";
    wire::json_escape(o, r.rid);
    o += chunk ? ",\"object\":\"chat.completion.chunk\"" : ",\"object\":\"chat.completion\"";
    o += ",\"created\":" + std::to_string(r.created) + ",\"model\":";
    wire::json_escape(o, r.model);
    o += ",\"choices\":[";
    if (!no_choices) {
      o += chunk ? "{\"index\":0,\"delta\":{" : "{\"index\":0,\"message\":{";
      bool role = !chunk || !r.sent_role;
      if (role) { o += "\"role\":\"assistant\""; r.sent_role = true; }
      if (!chunk || !text.empty()) { o += role ? ",\"content\":" : "\"content\":"; wire::json_escape(o, text); }
      o += "},\"logprobs\":null,\"finish_reason\":";
      if (finish) { o.push_back('"'); o += finish; o.push_back('"'); } else o += "null";
      o += "}";
    }
    o += "]";
    if (usage) {
      o += ",\"usage\":{\"prompt_tokens\":" + std::to_string(r.n_prompt) + ",\"completion_tokens\":" + std::to_string(r.out_ids.size()) +
           ",\"total_tokens\":" + std::to_string(r.n_prompt + r.out_ids.size()) + "}";
    }
    o += "}";
    return o;
  }

  // messages → prompt text in one of the built-in dialects (the Python server's entrypoints/chat_templates.py has the full
  // set and jinja; the daemon covers the two most common wire formats plus neutral role tags).  Returns false with `err`.
  bool render_chat(const Val& messages, bool add_generation_prompt, std::string& out, std::string& err) const {
    if (messages.t != Val::ARR || messages.a.empty()) { err = "`messages` must be a non-empty array"; return false; }
    const std::string& d = cfg.chat_template;
    if (d == "llama-3" || d == "llama3") out += "<|begin_of_text|>";
    for (auto& m : messages.a) {
      const Val* role = m.get("role");
      const Val* content = m.get("content");
      if (!role || role->t != Val::STR) { err = "every message needs a string `role`"; return false; }
      std::string text;
      if (content && content->t == Val::STR) text = content->s;
      else if (content && content->t == Val::ARR) {
        for (auto& part : content->a) {
          const Val* ty = part.get("type");
          const Val* tx = part.get("text");
          if (ty && ty->t == Val::STR && ty->s == "text" && tx && tx->t == Val::STR) { if (!text.empty()) text += "\n"; text += tx->s; }
          else { err = "the native front-end renders text content only; send multimodal messages to the Python API server"; return false; }
        }
      } else if (content && content->t != Val::NIL) { err = "`content` must be a string or an array of parts"; return false; }
      if (d == "chatml") out += "<|im_start|>" + role->s + "\n" + text + "<|im_end|>\n";
      else if (d == "llama-3" || d == "llama3") out += "<|start_header_id|>" + role->s + "<|end_header_id|>\n\n" + text + "<|eot_id|>";
      else out += "<|" + role->s + "|>\n" + text + "\n";
    }
    if (add_generation_prompt) {
      if (d == "chatml") out += "<|im_start|>assistant\n";
      else if (d == "llama-3" || d == "llama3") out += "<|start_header_id|>assistant<|end_header_id|>\n\n";
      else out += "<|assistant|>\n";
    }
    return true;
  }

  std::string completion_json(const Req& r, const std::string& text, const char* finish, bool chunk, bool usage,
                              bool no_choices = false) {
    std::string o = "{\"id\":";
    wire::json_escape(o, r.rid);
    o += ",\"object\":\"text_completion\",\"created\":" + std::to_string(r.created) + ",\"model\":";
    wire::json_escape(o, r.model);
    if (no_choices) o += ",\"choices\":[";
    else {
    o += ",\"choices\":[{\"index\":0,\"text\":";
    wire::json_escape(o, text);
    o += ",\"logprobs\":null,\"finish_reason\":";
    if (finish) { o.push_back('"'); o += finish; o.push_back('"'); } else o += "null";
    if (!chunk) {
      o += ",\"token_ids\":[";
      for (size_t k = 0; k < r.out_ids.size(); k++) { if (k) o.push_back(','); o += std::to_string(r.out_ids[k]); }
      o += "]";
    }
    o += "}";
    }
    o += "]";
    if (usage) {
      o += ",\"usage\":{\"prompt_tokens\":" + std::to_string(r.n_prompt) + ",\"completion_tokens\":" + std::to_string(r.out_ids.size()) +
           ",\"total_tokens\":" + std::to_string(r.n_prompt + r.out_ids.size()) + "}";
    }
    if (r.kv_params.t == Val::MAP) { o += ",\"kv_transfer_params\":"; wire::json_write(o, r.kv_params); }
    o += "}";
    return o;
  }

  std::string token_json(const Req& r, size_t from, const char* finish, bool final_nonstream) {
    std::string o = "{\"request_id\":";
    wire::json_escape(o, r.rid);
    o += ",\"token_ids\":[";
    for (size_t k = from; k < r.out_ids.size(); k++) { if (k > from) o.push_back(','); o += std::to_string(r.out_ids[k]); }
    o += "],\"finish_reason\":";
    if (finish) { o.push_back('"'); o += finish; o.push_back('"'); } else o += "null";
    if (final_nonstream) o += ",\"num_prompt_tokens\":" + std::to_string(r.n_prompt);
    o += "}";
    return o;
  }

  static std::string http_response(int code, const char* status, const std::string& body, const char* ctype = "application/json") {
    std::string o = "HTTP/1.1 " + std::to_string(code) + " " + status + "\r\nContent-Type: " + ctype +
                    "\r\nContent-Length: " + std::to_string(body.size()) + "\r\nConnection: keep-alive\r\n\r\n";
    return o + body;
  }
  static std::string sse_chunk(const std::string& payload) {
    std::string ev = "data: " + payload + "\n\n";
    char hdr[32]; snprintf(hdr, sizeof hdr, "%zx\r\n", ev.size());
    return std::string(hdr) + ev + "\r\n";
  }
  static std::string error_json(const std::string& msg, const char* type, int code) {
    std::string o = "{\"error\":{\"message\":";
    wire::json_escape(o, msg);
    o += ",\"type\":\""; o += type; o += "\",\"code\":" + std::to_string(code) + "}}";
    return o;
  }

  // ---------------------------------------------------------------- output thread
  void on_output(const Val& eco) {                      // mu held
    const Val* rid = eco.get("request_id");
    if (!rid) return;
    auto it = reqs.find(rid->s);
    if (it == reqs.end()) return;
    Req& r = *it->second;
    if (r.done) return;
    const Val* ids = eco.get("new_token_ids");
    size_t n_new = 0;
    if (ids && ids->t == Val::ARR) {
      for (auto& x : ids->a) {
        int64_t id = x.integer();
        r.out_ids.push_back(id); n_new++;
        if (r.detok && !r.token_api && id >= 0 && size_t(id) < vocab.bytes.size() && !(r.skip_special && vocab.special[id]))
          r.pending += vocab.bytes[id];
      }
    }
    n_gen_tokens += n_new;
    const Val* fin = eco.get("finish_reason");
    bool finished = fin && fin->t == Val::INT;
    const char* finish = finished ? finish_name(fin->i) : nullptr;
    if (const Val* kv = eco.get("kv_transfer_params")) r.kv_params = *kv;

    if (!r.token_api) {
      size_t scan_from = r.text.size();
      utf8_drain(r.pending, r.text, finished);
      for (auto& s : r.stops) {                          // stop strings (reference: detokenizer.py check_stop_strings)
        size_t from = scan_from > s.size() ? scan_from - s.size() + 1 : 0;
        size_t pos = r.text.find(s, from);
        if (pos != std::string::npos) {
          r.text.resize(r.include_stop ? pos + s.size() : pos);
          if (!finished) { aborts.push_back(r.rid); }
          finished = true; finish = "stop";
          break;
        }
      }
    }
    if (r.stream) {
      std::string data;
      if (r.token_api) {
        if (n_new || finished) data = sse_chunk(token_json(r, r.sent_ids, finish, false));
        r.sent_ids = r.out_ids.size();
      } else {
        size_t upto = finished ? r.text.size() : (r.text.size() > r.hold ? r.text.size() - r.hold : 0);
        while (!finished && upto > r.sent_text && (static_cast<unsigned char>(r.text[upto]) & 0xc0) == 0x80) upto--;   // never split a character
        if (upto > r.sent_text || finished) {
          std::string delta = upto > r.sent_text ? r.text.substr(r.sent_text, upto - r.sent_text) : std::string();
          if (upto > r.sent_text) r.sent_text = upto;
          if (!delta.empty() || finished)
            data = sse_chunk(r.chat ? chat_json(r, delta, finish, true, false) : completion_json(r, delta, finish, true, false));
        }
      }
      if (finished) {
        if (r.include_usage && !r.token_api) {
          data += sse_chunk(r.chat ? chat_json(r, "", nullptr, true, true, true) : completion_json(r, "", nullptr, true, true, true));
        }
        data += sse_chunk("[DONE]") + "0\r\n\r\n";
      }
      if (!data.empty()) outbox.push_back({r.conn_id, std::move(data), finished, false});
    } else if (finished) {
      std::string body = r.token_api ? token_json(r, 0, finish, true)
                         : r.chat ? chat_json(r, r.text, finish, false, true) : completion_json(r, r.text, finish, false, true);
      outbox.push_back({r.conn_id, http_response(200, "OK", body), true, false});
    }
    if (finished) {
      n_finished++;
      if (finish && !strcmp(finish, "abort")) n_aborted++;
      reqs.erase(it);
    }
  }

  void output_loop() {
    std::vector<std::string> fr;
    while (!stop) {
      if (!z::recv_multipart(pull, fr)) continue;
      if (fr.empty()) continue;
      if (fr[0] == "ENGINE_CORE_DEAD") {
        fprintf(stderr, "frontd: engine core died, shutting down\n");
        stop = true; wake(); return;
      }
      try {
        wire::MpReader rd(fr[0].data(), fr[0].size());
        Val top = rd.read();
        if (top.t != Val::ARR || top.a.size() < 2) continue;
        const Val& o = top.a[1];
        std::lock_guard<std::mutex> g(mu);
        if (const Val* st = o.get("scheduler_stats")) {
          if (const Val* x = st->get("num_running_reqs")) running = x->integer(); else running = 0;
          if (const Val* x = st->get("num_waiting_reqs")) waiting = x->integer(); else waiting = 0;
          if (const Val* x = st->get("kv_cache_usage")) kv_usage = x->num(); else kv_usage = 0;
        }
        if (const Val* outs = o.get("outputs")) for (auto& e : outs->a) on_output(e);
        if (!outbox.empty() || !aborts.empty()) wake();
      } catch (const std::exception& ex) {
        fprintf(stderr, "frontd: bad engine message: %s\n", ex.what());
      }
    }
  }

  // ---------------------------------------------------------------- HTTP side
  void conn_write(Conn& c) {
    while (!c.out.empty()) {
      ssize_t n = ::send(c.fd, c.out.data(), c.out.size(), MSG_NOSIGNAL);
      if (n > 0) { c.out.erase(0, n); continue; }
      if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) break;
      close_conn(c.id); return;
    }
    epoll_event ev{}; ev.data.fd = c.fd; ev.events = c.out.empty() ? uint32_t(EPOLLIN) : uint32_t(EPOLLIN | EPOLLOUT);
    epoll_ctl(ep, EPOLL_CTL_MOD, c.fd, &ev);
    if (c.out.empty() && c.close_after_flush && c.rid.empty()) close_conn(c.id);
  }

  void close_conn(uint64_t id) {
    auto it = conns.find(id);
    if (it == conns.end()) return;
    Conn& c = *it->second;
    if (!c.rid.empty()) {                                   // client went away mid-request → abort in the engine
      std::lock_guard<std::mutex> g(mu);
      auto r = reqs.find(c.rid);
      if (r != reqs.end()) { reqs.erase(r); send_abort(c.rid); n_aborted++; }
    }
    epoll_ctl(ep, EPOLL_CTL_DEL, c.fd, nullptr);
    ::close(c.fd);
    fd2conn.erase(c.fd);
    conns.erase(it);
  }

  void reply(Conn& c, int code, const char* status, const std::string& body, const char* ctype = "application/json") {
    c.out += http_response(code, status, body, ctype);
    conn_write(c);
  }

  static bool ieq(const std::string& a, const char* b) { return strcasecmp(a.c_str(), b) == 0; }

  // pulls SamplingParams fields out of a JSON object into a msgpack map (only what was given)
  static void sampling_fields(const Val& src, std::vector<std::pair<std::string, Val>>& dst) {
    static const char* keys[] = {"temperature", "top_p", "top_k", "min_p", "seed", "presence_penalty", "frequency_penalty",
                                 "repetition_penalty", "ignore_eos", "min_tokens", "stop_token_ids", "logit_bias",
                                 "allowed_token_ids", "bad_words_token_ids", "max_tokens"};
    for (auto k : keys) if (const Val* v = src.get(k)) if (v->t != Val::NIL) dst.emplace_back(k, *v);
  }

  void handle_generate(Conn& c, const Val& body, bool token_api, const std::string* chat_prompt = nullptr) {
    std::vector<int64_t> prompt;
    Val chat_val;
    if (chat_prompt) { chat_val.t = Val::STR; chat_val.s = *chat_prompt; }
    const Val* p = chat_prompt ? &chat_val : token_api ? body.get("token_ids") : body.get("prompt");
    if (token_api && !p) p = body.get("prompt_token_ids");
    if (p && p->t == Val::ARR && p->a.size() == 1 && (p->a[0].t == Val::ARR || p->a[0].t == Val::STR)) p = &p->a[0];
    if (p && p->t == Val::STR && vocab.enc.ready) {              // text prompt: native BPE encoder
      bool bos = chat_prompt == nullptr;                    // chat dialects spell their own begin-of-text marker
      if (const Val* v = body.get("add_special_tokens")) bos = v->truthy();
      for (auto t : vocab.enc.encode(p->s, bos)) prompt.push_back(t);
      if (prompt.empty()) { reply(c, 400, "Bad Request", error_json("the prompt is empty", "BadRequestError", 400)); return; }
    } else
    if (!p || p->t != Val::ARR || p->a.empty() || p->a[0].t != Val::INT) {
      reply(c, 400, "Bad Request", error_json("the native front-end takes token-id prompts (a non-empty array of integers); "
                                              "send text prompts to the Python API server or tokenise client-side", "BadRequestError", 400));
      return;
    }
    if (prompt.empty()) for (auto& x : p->a) prompt.push_back(x.integer());
    if (int64_t(prompt.size()) >= cfg.max_model_len) {
      reply(c, 400, "Bad Request", error_json("prompt is longer than the model's context window", "BadRequestError", 400));
      return;
    }
    std::vector<std::pair<std::string, Val>> sp;
    const Val* spsrc = token_api ? body.get("sampling_params") : &body;
    if (spsrc && spsrc->t == Val::MAP) sampling_fields(*spsrc, sp);
    int64_t max_tokens = 16;
    for (auto& kv : sp) if (kv.first == "max_tokens") max_tokens = kv.second.integer();
    if (const Val* v = body.get("max_completion_tokens")) if (v->t == Val::INT) max_tokens = v->i;
    max_tokens = std::max<int64_t>(1, std::min<int64_t>(max_tokens, cfg.max_model_len - int64_t(prompt.size())));

    auto r = std::make_unique<Req>();
    r->rid = std::string(token_api ? "gen-" : "cmpl-") + std::to_string(getpid()) + "-" + std::to_string(next_req++);
    if (const Val* v = body.get("request_id")) if (v->t == Val::STR) r->rid = "gen-" + v->s;
    r->conn_id = c.id; r->token_api = token_api; r->n_prompt = prompt.size(); r->created = int64_t(now_s());
    if (chat_prompt) { r->chat = true; r->rid = "chatcmpl-" + std::to_string(getpid()) + "-" + std::to_string(next_req - 1); }
    r->model = cfg.model;
    if (const Val* v = body.get("stream")) r->stream = v->truthy();
    if (const Val* v = body.get("stream_options")) if (const Val* u = v->get("include_usage")) r->include_usage = u->truthy();
    if (const Val* v = body.get("skip_special_tokens")) r->skip_special = v->truthy();
    if (const Val* v = body.get("include_stop_str_in_output")) r->include_stop = v->truthy();
    if (const Val* v = body.get("detokenize")) r->detok = v->truthy();
    if (const Val* v = body.get("stop")) {
      if (v->t == Val::STR) r->stops.push_back(v->s);
      else if (v->t == Val::ARR) for (auto& s : v->a) if (s.t == Val::STR && !s.s.empty()) r->stops.push_back(s.s);
    }
    if (!r->include_stop) for (auto& s : r->stops) r->hold = std::max(r->hold, s.size() - 1);
    if (vocab.bytes.empty()) r->detok = false;

    // EngineCoreRequest (msgspec map, defaults omitted)
    wire::MpWriter w;
    const Val* kvp = body.get("kv_transfer_params");
    const Val* prio = body.get("priority");
    const Val* salt = body.get("cache_salt");
    int nfields = 5 + (kvp && kvp->t == Val::MAP) + (prio && prio->t == Val::INT && prio->i) + (salt && salt->t == Val::STR) + (cfg.eos >= 0);
    w.map(nfields);
    w.str("request_id"); w.str(r->rid);
    w.str("prompt_token_ids"); w.array(prompt.size()); for (auto t : prompt) w.integer(t);
    w.str("sampling_params");
    {
      // stop ids: user's + EOS (unless ignore_eos) — what the Python input processor derives
      std::vector<int64_t> all_stop;
      bool ignore_eos = false;
      size_t n = 0;
      for (auto& kv : sp) { if (kv.first != "max_tokens") n++; if (kv.first == "ignore_eos") ignore_eos = kv.second.truthy();
                            if (kv.first == "stop_token_ids" && kv.second.t == Val::ARR) for (auto& x : kv.second.a) all_stop.push_back(x.integer()); }
      if (!ignore_eos && cfg.eos >= 0) all_stop.push_back(cfg.eos);
      w.map(n + 3 + (r->hold ? 1 : 0));
      for (auto& kv : sp) if (kv.first != "max_tokens") { w.str(kv.first); w.value(kv.second); }
      w.str("max_tokens"); w.integer(max_tokens);
      w.str("detokenize"); w.boolean(false);
      w.str("all_stop_token_ids"); w.array(all_stop.size()); for (auto t : all_stop) w.integer(t);
      if (r->hold) { w.str("output_text_buffer_length"); w.integer(int64_t(r->hold)); }
    }
    if (cfg.eos >= 0) { w.str("eos_token_id"); w.integer(cfg.eos); }
    w.str("arrival_time"); w.real(now_s());
    w.str("client_index"); w.integer(0);
    if (kvp && kvp->t == Val::MAP) { w.str("kv_transfer_params"); w.value(*kvp); }
    if (prio && prio->t == Val::INT && prio->i) { w.str("priority"); w.integer(prio->i); }
    if (salt && salt->t == Val::STR) { w.str("cache_salt"); w.str(salt->s); }

    c.rid = r->rid;
    bool stream = r->stream;
    n_requests++; n_prompt_tokens += prompt.size();
    {
      std::lock_guard<std::mutex> g(mu);
      reqs.emplace(r->rid, std::move(r));
    }
    if (stream) {
      c.streaming = true;
      c.out += "HTTP/1.1 200 OK\r\nContent-Type: text/event-stream\r\nCache-Control: no-cache\r\nTransfer-Encoding: chunked\r\nConnection: keep-alive\r\n\r\n";
      conn_write(c);
    }
    send_engine("A", w.o);
  }

  std::string metrics_text() {
    std::lock_guard<std::mutex> g(mu);
    char b[1536];
    snprintf(b, sizeof b,
             "# TYPE vllm:num_requests_running gauge\nvllm:num_requests_running{model_name=\"%s\"} %lld\n"
             "# TYPE vllm:num_requests_waiting gauge\nvllm:num_requests_waiting{model_name=\"%s\"} %lld\n"
             "# TYPE vllm:kv_cache_usage_perc gauge\nvllm:kv_cache_usage_perc{model_name=\"%s\"} %.6f\n"
             "# TYPE vllm:prompt_tokens_total counter\nvllm:prompt_tokens_total{model_name=\"%s\"} %llu\n"
             "# TYPE vllm:generation_tokens_total counter\nvllm:generation_tokens_total{model_name=\"%s\"} %llu\n"
             "# TYPE vllm:request_success_total counter\nvllm:request_success_total{model_name=\"%s\"} %llu\n"
             "# TYPE vllm:frontend_requests_total counter\nvllm:frontend_requests_total{model_name=\"%s\"} %llu\n"
             "# TYPE vllm:frontend_aborted_total counter\nvllm:frontend_aborted_total{model_name=\"%s\"} %llu\n",
             cfg.model.c_str(), (long long)running, cfg.model.c_str(), (long long)waiting, cfg.model.c_str(), kv_usage,
             cfg.model.c_str(), (unsigned long long)n_prompt_tokens.load(), cfg.model.c_str(), (unsigned long long)n_gen_tokens.load(),
             cfg.model.c_str(), (unsigned long long)n_finished.load(), cfg.model.c_str(), (unsigned long long)n_requests.load(),
             cfg.model.c_str(), (unsigned long long)n_aborted.load());
    return b;
  }

  void handle_http(Conn& c, const std::string& method, const std::string& path,
                   const std::map<std::string, std::string>& hdr, const std::string& body) {
    if (!cfg.api_key.empty() && path.rfind("/v1", 0) == 0) {
      auto it = hdr.find("authorization");
      if (it == hdr.end() || it->second != "Bearer " + cfg.api_key) { reply(c, 401, "Unauthorized", error_json("invalid API key", "AuthenticationError", 401)); return; }
    }
    if (method == "GET") {
      if (path == "/health" || path == "/ping") { reply(c, 200, "OK", "{}"); return; }
      if (path == "/version") { reply(c, 200, "OK", "{\"version\":\"native-frontd\"}"); return; }
      if (path == "/metrics") { reply(c, 200, "OK", metrics_text(), "text/plain; version=0.0.4"); return; }
      if (path == "/v1/models") {
        std::string o = "{\"object\":\"list\",\"data\":[{\"id\":";
        wire::json_escape(o, cfg.model);
        o += ",\"object\":\"model\",\"owned_by\":\"vllm_b200\",\"max_model_len\":" + std::to_string(cfg.max_model_len) + "}]}";
        reply(c, 200, "OK", o); return;
      }
      reply(c, 404, "Not Found", error_json("no such route", "NotFoundError", 404)); return;
    }
    if (method == "POST" && path == "/v1/chat/completions") {
      Val v;
      try { wire::JsonReader jr(body); v = jr.read(); } catch (const std::exception& e) { reply(c, 400, "Bad Request", error_json(e.what(), "BadRequestError", 400)); return; }
      if (v.t != Val::MAP) { reply(c, 400, "Bad Request", error_json("body must be a JSON object", "BadRequestError", 400)); return; }
      if (!vocab.enc.ready) { reply(c, 400, "Bad Request", error_json("chat completions need the native tokenizer (start the daemon with --vocab)", "BadRequestError", 400)); return; }
      if (v.get("tools") || v.get("response_format")) { reply(c, 400, "Bad Request", error_json("tool calling / structured output are served by the Python API server", "BadRequestError", 400)); return; }
      const Val* msgs = v.get("messages");
      bool gen = true;
      if (const Val* g = v.get("add_generation_prompt")) gen = g->truthy();
      std::string prompt, err;
      if (!msgs || !render_chat(*msgs, gen, prompt, err)) { reply(c, 400, "Bad Request", error_json(err.empty() ? "`messages` is required" : err, "BadRequestError", 400)); return; }
      handle_generate(c, v, false, &prompt);
      return;
    }
    if (method == "POST" && (path == "/v1/completions" || path == "/inference/v1/generate" || path == "/detokenize" || path == "/tokenize")) {
      Val v;
      try { wire::JsonReader jr(body); v = jr.read(); } catch (const std::exception& ex) { reply(c, 400, "Bad Request", error_json(ex.what(), "BadRequestError", 400)); return; }
      if (v.t != Val::MAP) { reply(c, 400, "Bad Request", error_json("body must be a JSON object", "BadRequestError", 400)); return; }
      if (path == "/tokenize") {
        const Val* pr = v.get("prompt");
        if (!vocab.enc.ready || !pr || pr->t != Val::STR) { reply(c, 400, "Bad Request", error_json("text tokenisation is not available for this tokenizer on the native tier", "BadRequestError", 400)); return; }
        bool bos = true;
        if (const Val* x = v.get("add_special_tokens")) bos = x->truthy();
        auto ids = vocab.enc.encode(pr->s, bos);
        std::string o = "{\"count\":" + std::to_string(ids.size()) + ",\"max_model_len\":" + std::to_string(cfg.max_model_len) + ",\"tokens\":[";
        for (size_t k = 0; k < ids.size(); k++) { if (k) o.push_back(','); o += std::to_string(ids[k]); }
        o += "]}";
        reply(c, 200, "OK", o); return;
      }
      if (path == "/detokenize") {
        std::string raw, text;
        if (const Val* t = v.get("tokens")) for (auto& x : t->a) { int64_t id = x.integer(); if (id >= 0 && size_t(id) < vocab.bytes.size() && !vocab.special[id]) raw += vocab.bytes[id]; }
        utf8_drain(raw, text, true);
        std::string o = "{\"prompt\":"; wire::json_escape(o, text); o += "}";
        reply(c, 200, "OK", o); return;
      }
      if (const Val* n = v.get("n")) if (n->t == Val::INT && n->i != 1) { reply(c, 400, "Bad Request", error_json("n > 1 is served by the Python API server", "BadRequestError", 400)); return; }
      handle_generate(c, v, path != "/v1/completions");
      return;
    }
    reply(c, 404, "Not Found", error_json("no such route", "NotFoundError", 404));
  }

  // returns false if more bytes are needed
  bool try_parse(Conn& c) {
    if (!c.rid.empty()) return false;                      // one request at a time per connection (no pipelining)
    size_t he = c.in.find("\r\n\r\n");
    if (he == std::string::npos) { if (c.in.size() > (1 << 20)) close_conn(c.id); return false; }
    size_t le = c.in.find("\r\n");
    std::string line = c.in.substr(0, le);
    size_t s1 = line.find(' '), s2 = line.rfind(' ');
    if (s1 == std::string::npos || s2 <= s1) { close_conn(c.id); return false; }
    std::string method = line.substr(0, s1), path = line.substr(s1 + 1, s2 - s1 - 1);
    size_t q = path.find('?'); if (q != std::string::npos) path.resize(q);
    std::map<std::string, std::string> hdr;
    size_t pos = le + 2;
    while (pos < he) {
      size_t eol = c.in.find("\r\n", pos);
      std::string h = c.in.substr(pos, eol - pos);
      size_t col = h.find(':');
      if (col != std::string::npos) {
        std::string k = h.substr(0, col), v = h.substr(col + 1);
        for (auto& ch : k) ch = char(tolower(ch));
        size_t a = v.find_first_not_of(' '); v = a == std::string::npos ? "" : v.substr(a);
        hdr[k] = v;
      }
      pos = eol + 2;
    }
    size_t clen = 0;
    if (hdr.count("content-length")) clen = strtoull(hdr["content-length"].c_str(), nullptr, 10);
    if (clen > (256u << 20)) { close_conn(c.id); return false; }
    if (c.in.size() < he + 4 + clen) return false;
    std::string body = c.in.substr(he + 4, clen);
    c.in.erase(0, he + 4 + clen);
    if (hdr.count("connection") && ieq(hdr["connection"], "close")) c.close_after_flush = true;
    uint64_t id = c.id;
    handle_http(c, method, path, hdr, body);
    return conns.count(id) != 0;
  }

  void drain_outbox() {
    std::deque<Outbox> q; std::deque<std::string> ab;
    { std::lock_guard<std::mutex> g(mu); q.swap(outbox); ab.swap(aborts); }
    for (auto& rid : ab) send_abort(rid);
    for (auto& o : q) {
      auto it = conns.find(o.conn_id);
      if (it == conns.end()) continue;
      Conn& c = *it->second;
      c.out += o.data;
      uint64_t id = c.id;
      if (o.end_of_response) { c.rid.clear(); c.streaming = false; }
      conn_write(c);
      if (o.end_of_response && conns.count(id)) while (try_parse(*conns[id])) {}
    }
  }

  void run() {
    z::load(cfg.libzmq.c_str());
    zctx = z::ctx_new();
    router = z::socket(zctx, z::ROUTER);
    pull = z::socket(zctx, z::PULL);
    int tmo = 500, linger = 0, mand = 0;
    z::setsockopt(router, z::RCVTIMEO, &tmo, sizeof tmo); z::setsockopt(pull, z::RCVTIMEO, &tmo, sizeof tmo);
    z::setsockopt(router, z::LINGER, &linger, sizeof linger); z::setsockopt(pull, z::LINGER, &linger, sizeof linger);
    z::setsockopt(router, z::ROUTER_MANDATORY, &mand, sizeof mand);
    if (z::bind(router, cfg.in_addr.c_str()) || z::bind(pull, cfg.out_addr.c_str())) {
      fprintf(stderr, "frontd: zmq bind failed: %s\n", z::strerror(z::zerrno())); exit(2);
    }
    if (!cfg.vocab_path.empty() && !vocab.load(cfg.vocab_path)) fprintf(stderr, "frontd: could not read %s; text output disabled\n", cfg.vocab_path.c_str());
    handshake();
    if (stop) return;

    lfd = ::socket(AF_INET, SOCK_STREAM | SOCK_NONBLOCK, 0);
    int one = 1; setsockopt(lfd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in sa{}; sa.sin_family = AF_INET; sa.sin_port = htons(cfg.port); inet_pton(AF_INET, cfg.host.c_str(), &sa.sin_addr);
    if (::bind(lfd, (sockaddr*)&sa, sizeof sa) || ::listen(lfd, 1024)) { perror("frontd: bind/listen"); exit(2); }
    ep = epoll_create1(0); evfd = eventfd(0, EFD_NONBLOCK);
    epoll_event ev{}; ev.events = EPOLLIN; ev.data.fd = lfd; epoll_ctl(ep, EPOLL_CTL_ADD, lfd, &ev);
    ev.data.fd = evfd; epoll_ctl(ep, EPOLL_CTL_ADD, evfd, &ev);
    std::thread out([this] { output_loop(); });
    fprintf(stderr, "frontd: serving %s on http://%s:%d\n", cfg.model.c_str(), cfg.host.c_str(), cfg.port);

    epoll_event evs[256];
    while (!stop) {
      int n = epoll_wait(ep, evs, 256, 200);
      for (int k = 0; k < n; k++) {
        int fd = evs[k].data.fd;
        if (fd == lfd) {
          while (true) {
            int cfd = accept4(lfd, nullptr, nullptr, SOCK_NONBLOCK);
            if (cfd < 0) break;
            setsockopt(cfd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
            auto c = std::make_unique<Conn>(); c->fd = cfd; c->id = next_conn++;
            epoll_event e2{}; e2.events = EPOLLIN; e2.data.fd = cfd; epoll_ctl(ep, EPOLL_CTL_ADD, cfd, &e2);
            fd2conn[cfd] = c->id; conns[c->id] = std::move(c);
          }
          continue;
        }
        if (fd == evfd) { uint64_t v; (void)!read(evfd, &v, 8); drain_outbox(); continue; }
        auto f = fd2conn.find(fd);
        if (f == fd2conn.end()) continue;
        uint64_t id = f->second;
        Conn& c = *conns[id];
        if (evs[k].events & (EPOLLHUP | EPOLLERR)) { close_conn(id); continue; }
        if (evs[k].events & EPOLLOUT) { conn_write(c); if (!conns.count(id)) continue; }
        if (evs[k].events & EPOLLIN) {
          char buf[65536]; bool dead = false;
          while (true) {
            ssize_t r = recv(fd, buf, sizeof buf, 0);
            if (r > 0) { c.in.append(buf, r); continue; }
            if (r == 0) dead = true;
            else if (errno != EAGAIN && errno != EWOULDBLOCK) dead = true;
            break;
          }
          if (dead) { close_conn(id); continue; }
          while (conns.count(id) && try_parse(*conns[id])) {}
        }
      }
      drain_outbox();
    }
    stop = true;
    out.join();
    for (auto& kv : conns) ::close(kv.second->fd);
    ::close(lfd);
    z::close(router); z::close(pull); z::ctx_term(zctx);
  }
};

static Server* g_srv = nullptr;
static void on_signal(int) { if (g_srv) { g_srv->stop = true; } }

int main(int argc, char** argv) {
  Server s;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string k = argv[i], v = argv[i + 1];
    if (k == "--host") s.cfg.host = v; else if (k == "--port") s.cfg.port = atoi(v.c_str());
    else if (k == "--in-addr") s.cfg.in_addr = v; else if (k == "--out-addr") s.cfg.out_addr = v;
    else if (k == "--model") s.cfg.model = v; else if (k == "--vocab") s.cfg.vocab_path = v;
    else if (k == "--chat-template") s.cfg.chat_template = v;
    else if (k == "--libzmq") s.cfg.libzmq = v; else if (k == "--api-key") s.cfg.api_key = v;
    else if (k == "--num-engines") s.cfg.num_engines = atoi(v.c_str());
    else if (k == "--max-model-len") s.cfg.max_model_len = atoll(v.c_str());
    else if (k == "--eos-token-id") s.cfg.eos = atoll(v.c_str());
    else { fprintf(stderr, "frontd: unknown option %s\n", k.c_str()); return 2; }
  }
  if (argc >= 2 && std::string(argv[argc - 1]) == "--encode-stdin") {
    // self-test / tooling: JSON array of strings on stdin → JSON array of id arrays on stdout
    if (!s.vocab.load(s.cfg.vocab_path) || !s.vocab.enc.ready) { fprintf(stderr, "frontd: no encoder in %s\n", s.cfg.vocab_path.c_str()); return 3; }
    std::string in; char buf[65536]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, stdin)) > 0) in.append(buf, n);
    wire::JsonReader jr(in);
    Val arr = jr.read();
    std::string o = "[";
    for (size_t k = 0; k < arr.a.size(); k++) {
      if (k) o.push_back(',');
      o.push_back('[');
      auto ids = s.vocab.enc.encode(arr.a[k].s, true);
      for (size_t j = 0; j < ids.size(); j++) { if (j) o.push_back(','); o += std::to_string(ids[j]); }
      o.push_back(']');
    }
    o += "]\n";
    fputs(o.c_str(), stdout);
    return 0;
  }
  if (s.cfg.in_addr.empty() || s.cfg.out_addr.empty() || s.cfg.libzmq.empty()) {
    fprintf(stderr, "usage: _frontd --in-addr ADDR --out-addr ADDR --libzmq PATH [--host H --port P --model NAME --vocab FILE "
                    "--max-model-len N --eos-token-id ID --num-engines N --api-key KEY]\n");
    return 2;
  }
  g_srv = &s;
  signal(SIGINT, on_signal); signal(SIGTERM, on_signal); signal(SIGPIPE, SIG_IGN);
  s.run();
  return 0;
}
