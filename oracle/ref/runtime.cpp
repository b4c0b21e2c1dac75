// TEST INFRASTRUCTURE — part of oracle/_ref/liblightmetrica.so (never shipped, never on the
// product path). Host-runtime shims the reference's hot-path sources link against when they are
// compiled without Boost/TBB/yaml-cpp:
//   * ComponentFactory_* : same registry semantics as /root/reference/src/liblightmetrica/component.cpp:56-125
//     (key -> {create, release}; duplicate key warns on stderr and overwrites; plugins are
//     dlopen()ed RTLD_LAZY|RTLD_LOCAL and register from their static initialisers).
//   * Logger_*           : stderr sink with the REAL argument order of logger.cpp:325
//     (type, message, filename, line, inplace, simple).
//   * PropertyUtils::PrintPrettyError : no-op (propertyutils.cpp:32-48 needs boost::format).
#include <pch.h>
#include <lightmetrica/component.h>
#include <lightmetrica/logger.h>
#include <lightmetrica/detail/propertyutils.h>
#include <dlfcn.h>

namespace lightmetrica_v2 {

namespace {
struct Funcs { CreateFuncPointerType create; ReleaseFuncPointerType release; };
std::unordered_map<std::string, Funcs>& Registry()
{
    static std::unordered_map<std::string, Funcs> m;
    return m;
}
std::vector<void*>& Plugins()
{
    static std::vector<void*> v;
    return v;
}
int g_verbose = 0;   // 0: errors+warnings only, 1: +info
}

extern "C" {

void ComponentFactory_Register(const char* key, CreateFuncPointerType c, ReleaseFuncPointerType r)
{
    auto& m = Registry();
    if (m.find(key) != m.end())
        std::cerr << "Failed to register [ " << key << " ]. Already registered." << std::endl;
    m[key] = Funcs{c, r};
}

void ComponentFactory_Unregister(const char* key) { Registry().erase(key); }

Component* ComponentFactory_Create(const char* key)
{
    auto it = Registry().find(key);
    if (it == Registry().end()) return nullptr;
    auto* p = it->second.create();
    p->createFunc = it->second.create;
    p->releaseFunc = it->second.release;
    p->createKey = it->first.c_str();   // stable storage (the reference stores the caller's pointer)
    return p;
}

ReleaseFuncPointerType ComponentFactory_ReleaseFunc(const char* key)
{
    auto it = Registry().find(key);
    return it == Registry().end() ? nullptr : it->second.release;
}

bool ComponentFactory_LoadPlugin(const char* path)
{
    const std::string p = std::string(path) + ".so";   // static.h:66-67 appends the extension
    void* h = dlopen(p.c_str(), RTLD_LAZY | RTLD_LOCAL);
    if (!h) {
        std::cerr << "Failed to load library or its dependencies : " << p << "\n" << dlerror() << std::endl;
        return false;
    }
    Plugins().push_back(h);
    return true;
}

void ComponentFactory_LoadPlugins(const char*) {}

void ComponentFactory_UnloadPlugins()
{
    for (void* h : Plugins()) dlclose(h);
    Plugins().clear();
}

void Logger_Run() {}
void Logger_Stop() {}
void Logger_SetVerboseLevel(int level) { g_verbose = level; }
void Logger_Log(int type, const char* message, const char* filename, int line, bool, bool)
{
    // LogType: Error=0, Warn=1, Info=2, Debug=3 (logger.h:45-51)
    if (type == 0 || (type == 1 && g_verbose >= 1) || (type >= 2 && g_verbose >= 2))
        fprintf(stderr, "[lm:%d] %s (%s:%d)\n", type, message, filename, line);
}
void Logger_UpdateIndentation(bool) {}
void Logger_Flush() {}

}  // extern "C"

void PropertyUtils::PrintPrettyError(const PropertyNode*) {}

}  // namespace lightmetrica_v2
