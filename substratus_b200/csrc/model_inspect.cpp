// model_inspect.cpp — ssb_model_read_tensor (include/ssb.h): host-only view of a Model artifact through the engine's
// own container readers.  No device, no arithmetic.
#include <cstring>

#include "../../include/ssb.h"
#include "loader.h"

namespace ssb {
void set_error(const std::string& s);  // engine.cu (thread-local text behind ssb_last_error)
}

extern "C" int ssb_model_read_tensor(const char* model_dir, const char* name, void* dst, int64_t cap, int64_t* nbytes_out, int* dtype,
                                     int64_t* shape, int* ndim) {
  if (!model_dir || !nbytes_out) {
    ssb::set_error("null argument");
    return SSB_EINVAL;
  }
  try {
    ssb::ModelFiles files;
    std::string err;
    if (!files.open(model_dir, &err)) {
      ssb::set_error(err);
      return SSB_EIO;
    }
    if (!name) {
      *nbytes_out = (int64_t)files.n_tensors();
      return SSB_OK;
    }
    const ssb::TensorView* tv = files.find(name);
    if (!tv) {
      ssb::set_error(std::string("no tensor named ") + name + " in " + model_dir);
      return SSB_EIO;
    }
    *nbytes_out = (int64_t)tv->nbytes;
    if (dtype) *dtype = tv->dtype;
    if (ndim) *ndim = (int)tv->shape.size();
    if (shape)
      for (size_t i = 0; i < tv->shape.size() && i < 4; ++i) shape[i] = tv->shape[i];
    if (dst || cap > 0) {
      if (!dst || cap < (int64_t)tv->nbytes) {
        ssb::set_error("destination buffer too small");
        return SSB_ENOMEM;
      }
      if (tv->nbytes) memcpy(dst, tv->data, tv->nbytes);
    }
    return SSB_OK;
  } catch (std::exception& ex) {
    ssb::set_error(std::string("exception: ") + ex.what());
    return SSB_EINVAL;
  }
}
