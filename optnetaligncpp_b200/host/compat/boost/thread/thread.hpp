// Stand-in for <boost/thread/thread.hpp> (absent in this image): nsgaiimain.cpp:9 includes it
// but uses nothing from it.
#pragma once
#include <thread>
