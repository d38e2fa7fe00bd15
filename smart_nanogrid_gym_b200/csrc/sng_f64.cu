// sng_f64.cu -- float64 validation build of the step (bit-faithful arithmetic: compiled with
// -fmad=false, one thread per env, numpy's summation order).  See DESIGN.md "Precision".
#include "sng_engine.cuh"

namespace sng {
EngineBase *make_engine_f64(const sng_config &cfg, int device, std::string &err)
{
    auto *e = new Engine<double, true>();
    if (e->init(cfg, device) != SNG_OK) {
        err = e->error;
        delete e;
        return nullptr;
    }
    return e;
}
}  // namespace sng
