// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
// Logging sink for the reference's formatted_output:: calls
// (include/formatted_output.h:21-29): silent unless CCAKE_REF_VERBOSE is set.
#include <cstdlib>
#include <iostream>
#include <string>
namespace formatted_output {
static bool verbose() { static const bool v = std::getenv("CCAKE_REF_VERBOSE") != nullptr; return v; }
void announce(std::string m) { if (verbose()) std::cout << "== " << m << std::endl; }
void summarize(std::string m) { if (verbose()) std::cout << m << std::endl; }
void report(std::string m) { if (verbose()) std::cout << "- " << m << std::endl; }
void update(std::string m) { if (verbose()) std::cout << "  * " << m << std::endl; }
void comment(std::string m) { if (verbose()) std::cout << "  * " << m << std::endl; }
void detail(std::string m) { if (verbose()) std::cout << "    - " << m << std::endl; }
}
