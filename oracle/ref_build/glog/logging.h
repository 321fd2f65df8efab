// Stand-in for the external google-glog header, which this image does not have.  The reference's matcher sources
// (okvis_matcher/include/okvis/ThreadPool.hpp:116) use exactly one glog facility, the LOG(ERROR) stream; this
// header maps it to std::cerr so that those sources compile unmodified from /root/reference (oracle/Makefile.ref).
// Test infrastructure only.
#pragma once
#include <iostream>
#define LOG(severity) (std::cerr << "[" #severity "] ")
