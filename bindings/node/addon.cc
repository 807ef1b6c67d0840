// Node-API addon over libzkcensus_b200: fullProve(inputs, wasmPath, zkeyPath) -> Promise<{proof, publicSignals}> with
// the shape of snarkjs groth16.fullProve (called at ts_inputs/src/example.ts:358-362 of the reference).  NOT compiled
// in this repository's image (no node / node_api.h); see INTEGRATION.md section 2.
// build: node-gyp with include_dirs = [<!(node -p "require('node-addon-api').include"), ../../include],
//        libraries = [-lzkcensus_b200, -lcudart]
#include <napi.h>
#include <fstream>
#include <map>
#include <vector>
#include "zkcensus_b200.h"

static zkb_ctx *g_ctx = nullptr;
static std::map<std::string, zkb_circuit *> g_circuits;   // key: zkeyPath + "|" + wasmPath

static std::vector<char> slurp(const std::string &p) {
  std::ifstream f(p, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), {});
}

class ProveWorker : public Napi::AsyncWorker {
 public:
  ProveWorker(Napi::Env env, std::string inputs, std::string wasm, std::string zkey)
      : Napi::AsyncWorker(env), deferred(Napi::Promise::Deferred::New(env)), inputs_(inputs), wasm_(wasm), zkey_(zkey) {}
  Napi::Promise::Deferred deferred;
  void Execute() override {
    if (!g_ctx && zkb_ctx_create(0, &g_ctx)) return SetError(zkb_last_error());
    zkb_circuit *&c = g_circuits[zkey_ + "|" + wasm_];
    if (!c) {
      auto z = slurp(zkey_), w = slurp(wasm_);
      if (zkb_load_circuit(g_ctx, z.data(), z.size(), w.data(), w.size(), &c)) return SetError(zkb_last_error());
    }
    proof_.resize(1024); pub_.resize(2048);
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

static Napi::Value FullProve(const Napi::CallbackInfo &info) {
  Napi::Env env = info.Env();
  auto stringify = env.Global().Get("JSON").As<Napi::Object>().Get("stringify").As<Napi::Function>();
  std::string inputs = info[0].IsString() ? info[0].As<Napi::String>().Utf8Value()
                                          : stringify.Call({info[0]}).As<Napi::String>().Utf8Value();
  auto *w = new ProveWorker(env, inputs, info[1].As<Napi::String>(), info[2].As<Napi::String>());
  w->Queue();
  return w->deferred.Promise();
}

Napi::Object Init(Napi::Env env, Napi::Object exports) {
  exports.Set("fullProve", Napi::Function::New(env, FullProve));
  return exports;
}
NODE_API_MODULE(zkcensus_b200, Init)
