// Node-API addon over libzkcensus_b200:
//   fullProve(inputs, wasmPath, zkeyPath) -> Promise<{proof, publicSignals}>   snarkjs groth16.fullProve
//       (called at ts_inputs/src/example.ts:358-362 of the reference)
//   verify(vKey, publicSignals, proof)    -> Promise<boolean>                  snarkjs groth16.verify
// NOT compiled in this repository's image (no node / node_api.h); see INTEGRATION.md section 2.
// build: node-gyp with include_dirs = [<!(node -p "require('node-addon-api').include"), ../../include],
//        libraries = [-lzkcensus_b200, -lcudart]
#include <napi.h>
#include <algorithm>
#include <fstream>
#include <map>
#include <mutex>
#include <vector>
#include "zkcensus_b200.h"

// Execute() runs on libuv worker threads, several at a time: the context and the circuit map are guarded by g_mu
// (held while a key is loaded, so two first calls for one key load it once); the proving call itself runs outside
// the lock - the library serialises calls per circuit handle.
static std::mutex g_mu;
static zkb_ctx *g_ctx = nullptr;
static std::map<std::string, zkb_circuit *> g_circuits;   // key: zkeyPath + "|" + wasmPath

static std::vector<char> slurp(const std::string &p) {
  std::ifstream f(p, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), {});
}

static zkb_circuit *circuit_for(const std::string &zkey, const std::string &wasm, std::string &err) {
  std::lock_guard<std::mutex> g(g_mu);
  if (!g_ctx && zkb_ctx_create(0, &g_ctx)) { err = zkb_last_error(); return nullptr; }
  auto it = g_circuits.find(zkey + "|" + wasm);
  if (it != g_circuits.end()) return it->second;
  auto z = slurp(zkey), w = slurp(wasm);
  if (z.empty() || w.empty()) { err = "cannot read " + (z.empty() ? zkey : wasm); return nullptr; }
  zkb_circuit *c = nullptr;
  if (zkb_load_circuit(g_ctx, z.data(), z.size(), w.data(), w.size(), &c)) { err = zkb_last_error(); return nullptr; }
  g_circuits[zkey + "|" + wasm] = c;
  return c;
}

class ProveWorker : public Napi::AsyncWorker {
 public:
  ProveWorker(Napi::Env env, std::string inputs, std::string wasm, std::string zkey)
      : Napi::AsyncWorker(env), deferred(Napi::Promise::Deferred::New(env)), inputs_(inputs), wasm_(wasm), zkey_(zkey) {}
  Napi::Promise::Deferred deferred;
  void Execute() override {
    std::string e;
    zkb_circuit *c = circuit_for(zkey_, wasm_, e);
    if (!c) return SetError(e);
    uint32_t info[8] = {0};
    zkb_circuit_info(c, info);
    proof_.resize(1024);
    pub_.resize(std::max<size_t>(2048, 96 * (size_t)info[1] + 64));   // public.json: <= 96 bytes per public signal
    size_t pn = proof_.size(), qn = pub_.size();
    char err[256] = {0};
    int rc = zkb_fullprove(c, inputs_.data(), inputs_.size(), &proof_[0], &pn, &pub_[0], &qn, err, sizeof err);
    if (rc) return SetError(err);
    proof_.resize(pn); pub_.resize(qn);
  }
  void OnOK() override {
    Napi::Env env = Env();
    auto JSONparse = env.Global().Get("JSON").As<Napi::Object>().Get("parse").As<Napi::Function>();
    Napi::Object proof = JSONparse.Call({Napi::String::New(env, proof_)}).As<Napi::Object>();
    proof.Set("protocol", "groth16");   // snarkjs adds these two fields
    proof.Set("curve", "bn128");
    Napi::Object out = Napi::Object::New(env);
    out.Set("proof", proof);
    out.Set("publicSignals", JSONparse.Call({Napi::String::New(env, pub_)}));
    deferred.Resolve(out);
  }
  void OnError(const Napi::Error &e) override { deferred.Reject(e.Value()); }
 private:
  std::string inputs_, wasm_, zkey_, proof_, pub_;
};

class VerifyWorker : public Napi::AsyncWorker {
 public:
  VerifyWorker(Napi::Env env, std::string vk, std::string pub, std::string proof)
      : Napi::AsyncWorker(env), deferred(Napi::Promise::Deferred::New(env)), vk_(vk), pub_(pub), proof_(proof) {}
  Napi::Promise::Deferred deferred;
  void Execute() override {
    int rc = zkb_verify(vk_.data(), vk_.size(), pub_.data(), pub_.size(), proof_.data(), proof_.size());
    if (rc == ZKB_OK) ok_ = true;
    else if (rc == ZKB_INVALID_PROOF) ok_ = false;      // snarkjs resolves false for an invalid proof
    else SetError(zkb_last_error());
  }
  void OnOK() override { deferred.Resolve(Napi::Boolean::New(Env(), ok_)); }
  void OnError(const Napi::Error &e) override { deferred.Reject(e.Value()); }
 private:
  std::string vk_, pub_, proof_;
  bool ok_ = false;
};

static std::string as_json(const Napi::CallbackInfo &info, size_t i) {
  Napi::Env env = info.Env();
  if (info[i].IsString()) return info[i].As<Napi::String>().Utf8Value();
  auto stringify = env.Global().Get("JSON").As<Napi::Object>().Get("stringify").As<Napi::Function>();
  return stringify.Call({info[i]}).As<Napi::String>().Utf8Value();
}

static Napi::Value FullProve(const Napi::CallbackInfo &info) {
  Napi::Env env = info.Env();
  if (info.Length() < 3 || !info[1].IsString() || !info[2].IsString()) {
    Napi::TypeError::New(env, "fullProve(inputs, wasmPath, zkeyPath)").ThrowAsJavaScriptException();
    return env.Undefined();
  }
  auto *w = new ProveWorker(env, as_json(info, 0), info[1].As<Napi::String>(), info[2].As<Napi::String>());
  w->Queue();
  return w->deferred.Promise();
}

// groth16.verify(vKey, publicSignals, proof): objects (as snarkjs takes them) or JSON strings
static Napi::Value Verify(const Napi::CallbackInfo &info) {
  Napi::Env env = info.Env();
  if (info.Length() < 3) {
    Napi::TypeError::New(env, "verify(vKey, publicSignals, proof)").ThrowAsJavaScriptException();
    return env.Undefined();
  }
  auto *w = new VerifyWorker(env, as_json(info, 0), as_json(info, 1), as_json(info, 2));
  w->Queue();
  return w->deferred.Promise();
}

Napi::Object Init(Napi::Env env, Napi::Object exports) {
  exports.Set("fullProve", Napi::Function::New(env, FullProve));
  exports.Set("verify", Napi::Function::New(env, Verify));
  return exports;
}
NODE_API_MODULE(zkcensus_b200, Init)
