// Stand-in for the file the reference's configure step generates
// (src/gitrev -> src/util/gitgen.cpp): three strings read by src/util/git.cpp:6-8.
const char* kGitRev = "unknown";
const char* kGitMsg = "";
const char* kGitDiff = "";
