// cvshim: stand-in for DBoW2/BowVector.h (DBoW2 is an un-vendored dependency of hySLAM, src/CMakeLists.txt:32).
// DBoW2 publishes BowVector as a std::map<WordId, WordValue>; only the type is needed for the hySLAM headers to parse.
#pragma once
#include <map>
namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {};
}
