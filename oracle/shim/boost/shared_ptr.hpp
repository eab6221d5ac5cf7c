// oracle shim (test infrastructure): Boost is absent in this image; the reference only
// uses boost::shared_ptr as a plain owning handle, so alias it to the std one.
#pragma once
#include <memory>
namespace boost { using std::shared_ptr; }
