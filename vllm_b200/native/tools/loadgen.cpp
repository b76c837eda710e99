// vllm_b200-loadgen: open-loop HTTP load generator for OpenAI-compatible completion servers.
//
// Native counterpart of the reference's Rust `bench` crate (rust/src/bench): one thread, non-blocking sockets +
// epoll, thousands of concurrent streaming requests without a Python interpreter in the measurement path.
// Every request POSTs /v1/completions with `stream:true`, token-id prompts of a fixed length and `ignore_eos`;
// time-to-first-token, inter-token latencies and end-to-end latency are stamped at the moment the bytes of an SSE
// event are read.  Arrivals follow a Poisson process (`--rate`, 0 = all at once), bounded by `--concurrency`.
// Output: one JSON object on stdout (same keys as `vllm_b200 bench serve`).
//
//   loadgen --host 127.0.0.1 --port 8000 --model llama-3-8b --num 1000 --input-len 1024 --output-len 128 --rate 20
#include <algorithm>
#include <arpa/inet.h>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <random>
#include <string>
#include <sys/epoll.h>
#include <sys/socket.h>
#include <unistd.h>
#include <vector>

using Clock = std::chrono::steady_clock;
static double now_s() { return std::chrono::duration<double>(Clock::now().time_since_epoch()).count(); }

struct Req {
  int fd = -1;
  std::string out;            // request bytes still to send
  size_t sent = 0;
  std::string in;             // unparsed response bytes
  bool headers_done = false;
  int status = 0;
  double t_start = 0, t_first = 0, t_last = 0, t_end = 0;
  std::vector<double> itl;
  int events = 0;             // SSE events carrying tokens
  int completion_tokens = -1; // from the usage block if present
  bool done = false, ok = true;
};

struct Args {
  std::string host = "127.0.0.1", model = "model", api_key;
  int port = 8000, num = 100, input_len = 128, output_len = 32, concurrency = 0, vocab = 32000;
  double rate = 0.0;
  uint64_t seed = 0;
};

static std::string build_request(const Args& a, std::mt19937_64& rng) {
  std::string body = "{\"model\":\"" + a.model + "\",\"prompt\":[";
  std::uniform_int_distribution<int> tok(5, std::max(a.vocab - 1, 6));
  for (int i = 0; i < a.input_len; ++i) { body += std::to_string(tok(rng)); if (i + 1 < a.input_len) body += ','; }
  body += "],\"max_tokens\":" + std::to_string(a.output_len) +
          ",\"temperature\":0.0,\"ignore_eos\":true,\"stream\":true,\"stream_options\":{\"include_usage\":true}}";
  std::string req = "POST /v1/completions HTTP/1.1\r\nHost: " + a.host + ":" + std::to_string(a.port) +
                    "\r\nContent-Type: application/json\r\nAccept: text/event-stream\r\nConnection: close\r\n";
  if (!a.api_key.empty()) req += "Authorization: Bearer " + a.api_key + "\r\n";
  req += "Content-Length: " + std::to_string(body.size()) + "\r\n\r\n" + body;
  return req;
}

static int connect_nb(const Args& a) {
  int fd = ::socket(AF_INET, SOCK_STREAM | SOCK_NONBLOCK, 0);
  if (fd < 0) return -1;
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  sockaddr_in sa{};
  sa.sin_family = AF_INET;
  sa.sin_port = htons((uint16_t)a.port);
  if (inet_pton(AF_INET, a.host.c_str(), &sa.sin_addr) != 1) { ::close(fd); return -1; }
  if (::connect(fd, (sockaddr*)&sa, sizeof(sa)) < 0 && errno != EINPROGRESS) { ::close(fd); return -1; }
  return fd;
}

// consume complete SSE events ("data: ...\n\n") from r.in; chunked transfer framing lines are ignored because an
// event never spans a line that looks like a hex chunk size AND contains "data:"
static void parse_events(Req& r, double t) {
  if (!r.headers_done) {
    size_t h = r.in.find("\r\n\r\n");
    if (h == std::string::npos) return;
    if (r.in.size() >= 12) r.status = atoi(r.in.c_str() + 9);
    r.in.erase(0, h + 4);
    r.headers_done = true;
    if (r.status != 200) { r.ok = false; r.done = true; return; }
  }
  for (;;) {
    size_t d = r.in.find("data:");
    if (d == std::string::npos) { if (r.in.size() > 16) r.in.erase(0, r.in.size() - 16); return; }
    size_t e = r.in.find("\n\n", d);
    size_t e2 = r.in.find("\r\n\r\n", d);
    if (e == std::string::npos && e2 == std::string::npos) { r.in.erase(0, d); return; }
    size_t end = std::min(e == std::string::npos ? r.in.size() : e, e2 == std::string::npos ? r.in.size() : e2);
    std::string ev = r.in.substr(d + 5, end - d - 5);
    r.in.erase(0, end + 2);
    if (ev.find("[DONE]") != std::string::npos) { r.done = true; r.t_end = t; return; }
    size_t u = ev.find("\"completion_tokens\":");
    if (u != std::string::npos) r.completion_tokens = atoi(ev.c_str() + u + 20);
    const bool has_tok = ev.find("\"text\":\"\"") == std::string::npos && ev.find("\"choices\":[]") == std::string::npos &&
                         ev.find("\"choices\"") != std::string::npos;
    if (has_tok) {
      if (r.events == 0) r.t_first = t; else r.itl.push_back(t - r.t_last);
      r.t_last = t;
      r.events++;
    }
  }
}

static double pct(std::vector<double>& v, double p) {
  if (v.empty()) return NAN;
  std::sort(v.begin(), v.end());
  double idx = p / 100.0 * (v.size() - 1);
  size_t lo = (size_t)idx;
  size_t hi = std::min(lo + 1, v.size() - 1);
  return v[lo] + (v[hi] - v[lo]) * (idx - lo);
}
static void stats(const char* name, std::vector<double> v, bool last) {
  double mean = 0;
  for (double x : v) mean += x;
  mean = v.empty() ? NAN : mean / v.size();
  printf("\"%s\":{\"mean\":%.4f,\"p50\":%.4f,\"p90\":%.4f,\"p99\":%.4f}%s", name, mean * 1e3, pct(v, 50) * 1e3, pct(v, 90) * 1e3,
         pct(v, 99) * 1e3, last ? "" : ",");
}

int main(int argc, char** argv) {
  Args a;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string k = argv[i], v = argv[i + 1];
    if (k == "--host") a.host = v; else if (k == "--port") a.port = atoi(v.c_str());
    else if (k == "--model") a.model = v; else if (k == "--num") a.num = atoi(v.c_str());
    else if (k == "--input-len") a.input_len = atoi(v.c_str()); else if (k == "--output-len") a.output_len = atoi(v.c_str());
    else if (k == "--rate") a.rate = atof(v.c_str()); else if (k == "--concurrency") a.concurrency = atoi(v.c_str());
    else if (k == "--vocab") a.vocab = atoi(v.c_str()); else if (k == "--seed") a.seed = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--api-key") a.api_key = v;
    else { fprintf(stderr, "unknown flag %s\n", k.c_str()); return 2; }
  }
  std::mt19937_64 rng(a.seed);
  std::exponential_distribution<double> gap(a.rate > 0 ? a.rate : 1.0);
  std::vector<Req> reqs(a.num);
  std::vector<double> arrival(a.num, 0.0);
  if (a.rate > 0) { double t = 0; for (int i = 0; i < a.num; ++i) { arrival[i] = t; t += gap(rng); } }
  int ep = epoll_create1(0);
  int next = 0, active = 0, finished = 0;
  const int max_active = a.concurrency > 0 ? a.concurrency : a.num;
  const double t0 = now_s();
  std::vector<epoll_event> evs(256);
  char buf[65536];
  while (finished < a.num) {
    const double t = now_s() - t0;
    while (next < a.num && active < max_active && arrival[next] <= t) {
      Req& r = reqs[next];
      r.out = build_request(a, rng);
      r.fd = connect_nb(a);
      r.t_start = now_s();
      if (r.fd < 0) { r.ok = false; r.done = true; finished++; next++; continue; }
      epoll_event e{};
      e.events = EPOLLIN | EPOLLOUT;
      e.data.u32 = (uint32_t)next;
      epoll_ctl(ep, EPOLL_CTL_ADD, r.fd, &e);
      active++; next++;
    }
    int timeout_ms = 50;
    if (next < a.num && active < max_active) timeout_ms = std::max(0, (int)((arrival[next] - (now_s() - t0)) * 1e3));
    int n = epoll_wait(ep, evs.data(), (int)evs.size(), std::min(timeout_ms, 50));
    for (int i = 0; i < n; ++i) {
      Req& r = reqs[evs[i].data.u32];
      if (r.fd < 0) continue;
      bool close_it = false;
      if ((evs[i].events & EPOLLOUT) && r.sent < r.out.size()) {
        ssize_t w = ::send(r.fd, r.out.data() + r.sent, r.out.size() - r.sent, MSG_NOSIGNAL);
        if (w > 0) r.sent += (size_t)w;
        else if (w < 0 && errno != EAGAIN && errno != EINPROGRESS && errno != ENOTCONN) { r.ok = false; close_it = true; }
        if (r.sent == r.out.size()) {
          epoll_event e{};
          e.events = EPOLLIN;
          e.data.u32 = evs[i].data.u32;
          epoll_ctl(ep, EPOLL_CTL_MOD, r.fd, &e);
        }
      }
      if (evs[i].events & (EPOLLIN | EPOLLHUP | EPOLLERR)) {
        for (;;) {
          ssize_t got = ::recv(r.fd, buf, sizeof(buf), 0);
          if (got > 0) { r.in.append(buf, (size_t)got); parse_events(r, now_s()); if (r.done) { close_it = true; break; } }
          else if (got == 0) { close_it = true; break; }
          else { if (errno != EAGAIN) { r.ok = false; close_it = true; } break; }
        }
      }
      if (close_it) {
        if (!r.done) { r.done = true; r.ok = r.ok && r.events > 0; }
        if (r.t_end == 0) r.t_end = now_s();
        epoll_ctl(ep, EPOLL_CTL_DEL, r.fd, nullptr);
        ::close(r.fd);
        r.fd = -1;
        active--; finished++;
      }
    }
  }
  const double wall = now_s() - t0;
  std::vector<double> ttft, e2e, tpot, itl;
  long out_tok = 0, in_tok = 0;
  int good = 0;
  for (auto& r : reqs) {
    if (!r.ok || r.events == 0) continue;
    good++;
    const int ntok = r.completion_tokens > 0 ? r.completion_tokens : r.events;
    out_tok += ntok; in_tok += a.input_len;
    ttft.push_back(r.t_first - r.t_start);
    e2e.push_back(r.t_end - r.t_start);
    if (ntok > 1) tpot.push_back((r.t_end - r.t_first) / (ntok - 1));
    itl.insert(itl.end(), r.itl.begin(), r.itl.end());
  }
  printf("{\"completed\":%d,\"failed\":%d,\"duration_s\":%.4f,\"request_throughput\":%.4f,\"output_throughput\":%.4f,"
         "\"total_token_throughput\":%.4f,", good, a.num - good, wall, good / wall, out_tok / wall, (in_tok + out_tok) / wall);
  stats("ttft_ms", ttft, false); stats("tpot_ms", tpot, false); stats("itl_ms", itl, false); stats("e2el_ms", e2e, true);
  printf("}\n");
  return good == a.num ? 0 : 1;
}
