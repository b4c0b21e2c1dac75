// TEST INFRASTRUCTURE — part of oracle/_ref/liblightmetrica.so (never shipped, never on the
// product path). Host services of Lightmetrica that are NOT on the hot path and whose real
// implementations need yaml-cpp / TBB / FreeImage / Boost, none of which exist here. Each shim
// keeps the contract of the file it stands in for:
//   PropertyTree_/Node  : /root/reference/src/liblightmetrica/property.cpp:40-230 (YAML block-style
//                         subset: nested maps, "- " sequences, one-line scalars; no flow style)
//   Assets_             : assets.cpp:50-122 (lazy create "<interface>::<type>" + Load(params))
//   film::hdr           : asset/film/film_hdr.cpp:193-305 minus FreeImage Save (writes .pfm)
//   Scheduler_          : scheduler.cpp:78-295 (std::thread instead of TBB; grain 10000; per-thread
//                         Film clone + Random seeded by initRng->NextUInt(); Clear/Accumulate/
//                         Rescale(W*H/N) at the end)
//   trianglemesh::mem   : an in-memory TriangleMesh so big synthetic meshes need no text parsing
#include <pch.h>
#include <lightmetrica/lightmetrica.h>
#include "host_shims.h"

LM_NAMESPACE_BEGIN

// ------------------------------------------------------------------------------------------------
// PropertyNode / PropertyTree (YAML block-style subset)

class TreeShim;

class NodeShim final : public PropertyNode
{
public:
    LM_IMPL_CLASS(NodeShim, PropertyNode);
    LM_IMPL_F(Tree)      = [this]() -> const PropertyTree* { return tree; };
    LM_IMPL_F(Type)      = [this]() -> PropertyNodeType { return type; };
    LM_IMPL_F(Line)      = [this]() -> int { return line; };
    LM_IMPL_F(Key)       = [this]() -> std::string { return key; };
    LM_IMPL_F(RawScalar) = [this]() -> const char* { return scalar.c_str(); };
    LM_IMPL_F(Size)      = [this]() -> int { return (int)seq.size(); };
    LM_IMPL_F(Child)     = [this](const std::string& k) -> const PropertyNode* { auto it = map.find(k); return it == map.end() ? nullptr : it->second; };
    LM_IMPL_F(At)        = [this](int i) -> const PropertyNode* { return seq.at(i); };
    LM_IMPL_F(Parent)    = [this]() -> const PropertyNode* { return parent; };
public:
    const PropertyTree* tree = nullptr;
    const NodeShim* parent = nullptr;
    PropertyNodeType type = PropertyNodeType::Scalar;
    int line = 0;
    std::string key, scalar;
    std::map<std::string, NodeShim*> map;
    std::vector<NodeShim*> seq;
};

class TreeShim final : public PropertyTree
{
public:
    LM_IMPL_CLASS(TreeShim, PropertyTree);
    LM_IMPL_F(LoadFromFile) = [this](const std::string& path) -> bool
    {
        std::ifstream f(path);
        if (!f) return false;
        std::stringstream ss; ss << f.rdbuf();
        path_ = path;
        return Parse(ss.str());
    };
    LM_IMPL_F(LoadFromString) = [this](const std::string& s) -> bool { return Parse(s); };
    LM_IMPL_F(LoadFromStringWithFilename) = [this](const std::string& s, const std::string& p, const std::string& bp) -> bool { path_ = p; basepath_ = bp; return Parse(s); };
    LM_IMPL_F(Path)     = [this]() -> std::string { return path_; };
    LM_IMPL_F(BasePath) = [this]() -> std::string { return basepath_; };
    LM_IMPL_F(Root)     = [this]() -> const PropertyNode* { return root_; };
    LM_IMPL_F(RawInput) = [this]() -> std::string { return input_; };

private:
    struct Line { int indent; std::string text; int no; };

    NodeShim* NewNode(const NodeShim* parent, int line)
    {
        pool_.emplace_back(new NodeShim);
        auto* n = pool_.back().get();
        n->tree = this; n->parent = parent; n->line = line;
        return n;
    }

    static std::string Trim(const std::string& s)
    {
        size_t b = s.find_first_not_of(" \t\r"), e = s.find_last_not_of(" \t\r");
        return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
    }

    // Parses the block starting at lines_[i] whose indentation is exactly `indent`.
    NodeShim* ParseBlock(size_t& i, int indent, const NodeShim* parent)
    {
        auto* node = NewNode(parent, lines_[i].no);
        const bool isSeq = lines_[i].text.compare(0, 2, "- ") == 0 || lines_[i].text == "-";
        node->type = isSeq ? PropertyNodeType::Sequence : PropertyNodeType::Map;
        while (i < lines_.size() && lines_[i].indent == indent)
        {
            if (isSeq)
            {
                if (!(lines_[i].text.compare(0, 2, "- ") == 0 || lines_[i].text == "-")) break;
                // Re-interpret "- rest" as a nested block indented by two more columns
                std::string rest = lines_[i].text.size() > 2 ? Trim(lines_[i].text.substr(2)) : std::string();
                if (rest.empty()) { i++; node->seq.push_back(ParseBlock(i, lines_[i].indent, node)); continue; }
                if (rest.find(": ") == std::string::npos && rest.back() != ':')
                {
                    auto* s = NewNode(node, lines_[i].no);
                    s->scalar = rest;
                    node->seq.push_back(s);
                    i++;
                    continue;
                }
                lines_[i].indent = indent + 2;
                lines_[i].text = rest;
                node->seq.push_back(ParseBlock(i, indent + 2, node));
            }
            else
            {
                const std::string& t = lines_[i].text;
                size_t c = t.find(':');
                if (c == std::string::npos) throw std::runtime_error("yaml-subset: expected 'key:' at line " + std::to_string(lines_[i].no));
                std::string key = Trim(t.substr(0, c));
                std::string val = Trim(t.substr(c + 1));
                if (!val.empty())
                {
                    auto* s = NewNode(node, lines_[i].no);
                    s->key = key; s->scalar = val;
                    node->map[key] = s;
                    i++;
                }
                else
                {
                    i++;
                    if (i < lines_.size() && lines_[i].indent > indent)
                    {
                        auto* ch = ParseBlock(i, lines_[i].indent, node);
                        ch->key = key;
                        node->map[key] = ch;
                    }
                    else
                    {
                        auto* s = NewNode(node, lines_[i - 1].no);
                        s->key = key;
                        node->map[key] = s;
                    }
                }
            }
        }
        return node;
    }

    bool Parse(const std::string& s)
    {
        input_ = s; lines_.clear(); pool_.clear(); root_ = nullptr;
        std::stringstream ss(s);
        std::string l; int no = 0;
        while (std::getline(ss, l))
        {
            no++;
            size_t h = l.find(" #");
            if (!l.empty() && l[0] == '#') continue;
            if (h != std::string::npos) l = l.substr(0, h);
            size_t b = l.find_first_not_of(' ');
            if (b == std::string::npos) continue;
            std::string t = Trim(l);
            if (t.empty()) continue;
            lines_.push_back(Line{(int)b, t, no});
        }
        if (lines_.empty()) return false;
        try { size_t i = 0; root_ = ParseBlock(i, lines_[0].indent, nullptr); }
        catch (const std::exception& e) { std::cerr << e.what() << std::endl; return false; }
        return true;
    }

    std::string path_, basepath_, input_;
    std::vector<Line> lines_;
    std::vector<std::unique_ptr<NodeShim>> pool_;
    NodeShim* root_ = nullptr;
};

LM_COMPONENT_REGISTER_IMPL(TreeShim, "PropertyTree_");

// ------------------------------------------------------------------------------------------------
// Assets (lazy loading; assets.cpp:50-122)

class AssetsShim final : public Assets
{
public:
    LM_IMPL_CLASS(AssetsShim, Assets);
    LM_IMPL_F(Initialize) = [this](const PropertyNode* p) -> bool { prop_ = p; return true; };
    LM_IMPL_F(AssetByIDAndType) = [this](const std::string& id, const std::string& iface, const Primitive* prim) -> Asset*
    {
        auto it = index_.find(id);
        if (it != index_.end()) return assets_[it->second].get();
        const auto* n = prop_ ? prop_->Child(id) : nullptr;
        if (!n) { LM_LOG_ERROR("Missing asset: " + id); return nullptr; }
        const auto* in = n->Child("interface");
        const auto* tn = n->Child("type");
        if (!in || !tn) { LM_LOG_ERROR("Asset needs 'interface' and 'type': " + id); return nullptr; }
        if (iface != in->RawScalar()) { LM_LOG_ERROR("Invalid asset interface for " + id); return nullptr; }
        auto a = ComponentFactory::Create<Asset>(iface + "::" + tn->RawScalar());
        if (!a) return nullptr;
        a->SetID(id);
        a->SetIndex((int)assets_.size());
        if (!a->Load(n->Child("params"), this, prim)) { LM_LOG_ERROR("Failed to load asset: " + id); return nullptr; }
        assets_.push_back(std::move(a));
        index_[id] = assets_.size() - 1;
        return assets_.back().get();
    };
    LM_IMPL_F(PostLoad) = [this](const Scene* scene) -> bool
    {
        for (auto& a : assets_) { if (a->PostLoad.Implemented() && !a->PostLoad(scene)) return false; }
        return true;
    };
    LM_IMPL_F(GetByIndex) = [this](int i) -> Asset* { return assets_.at(i).get(); };
private:
    const PropertyNode* prop_ = nullptr;
    std::vector<Asset::UniquePtr> assets_;
    std::unordered_map<std::string, size_t> index_;
};

LM_COMPONENT_REGISTER_IMPL(AssetsShim, "Assets_");

// ------------------------------------------------------------------------------------------------
// film::hdr without FreeImage (film_hdr.cpp:193-305). Save writes "<path>.pfm".

class FilmShim final : public Film
{
public:
    LM_IMPL_CLASS(FilmShim, Film);
    LM_IMPL_F(Load) = [this](const PropertyNode* p, Assets*, const Primitive*) -> bool
    {
        if (!p->ChildAs<int>("w", w_)) return false;
        if (!p->ChildAs<int>("h", h_)) return false;
        data_.assign((size_t)w_ * h_, Vec3());
        return true;
    };
    LM_IMPL_F(Clone) = [this](BasicComponent* o) -> void { auto* f = static_cast<FilmShim*>(o); f->w_ = w_; f->h_ = h_; f->data_ = data_; };
    LM_IMPL_F(Width)  = [this]() -> int { return w_; };
    LM_IMPL_F(Height) = [this]() -> int { return h_; };
    LM_IMPL_F(Splat) = [this](const Vec2& r, const SPD& v) -> void
    {
        const int px = Math::Clamp((int)(r.x * Float(w_)), 0, w_ - 1);
        const int py = Math::Clamp((int)(r.y * Float(h_)), 0, h_ - 1);
        data_[(size_t)py * w_ + px] += v.ToRGB();
    };
    LM_IMPL_F(SetPixel) = [this](int x, int y, const SPD& v) -> void { data_[(size_t)y * w_ + x] = v.ToRGB(); };
    LM_IMPL_F(Save) = [this](const std::string& path) -> bool
    {
        FILE* f = fopen((path + ".pfm").c_str(), "wb");
        if (!f) return false;
        fprintf(f, "PF\n%d %d\n-1.0\n", w_, h_);
        for (auto& v : data_) { float c[3] = {v.x, v.y, v.z}; fwrite(c, 4, 3, f); }
        fclose(f);
        return true;
    };
    LM_IMPL_F(Accumulate) = [this](const Film* o) -> void
    {
        auto* f = static_cast<const FilmShim*>(o);
        for (size_t i = 0; i < data_.size(); i++) data_[i] += f->data_[i];
    };
    LM_IMPL_F(Rescale) = [this](Float s) -> void { for (auto& v : data_) v *= s; };
    LM_IMPL_F(Clear) = [this]() -> void { data_.assign((size_t)w_ * h_, Vec3()); };
    LM_IMPL_F(PixelIndex) = [this](const Vec2& r) -> int
    {
        const int px = Math::Clamp((int)(r.x * Float(w_)), 0, w_ - 1);
        const int py = Math::Clamp((int)(r.y * Float(h_)), 0, h_ - 1);
        return py * w_ + px;
    };
public:
    int w_ = 0, h_ = 0;
    std::vector<Vec3> data_;
};

LM_COMPONENT_REGISTER_IMPL(FilmShim, "film::hdr");

auto RefHost::FilmData(const Film* film, int& w, int& h) -> const float*
{
    auto* f = static_cast<const FilmShim*>(film);
    w = f->w_; h = f->h_;
    return reinterpret_cast<const float*>(f->data_.data());   // Vec3 = 16 B (x,y,z,pad)
}

// ------------------------------------------------------------------------------------------------
// Scheduler_ (scheduler.cpp:78-295) with std::thread instead of TBB

int RefHost::numThreads = 1;

class SchedShim final : public Scheduler
{
public:
    LM_IMPL_CLASS(SchedShim, Scheduler);
    LM_IMPL_F(Load) = [this](const PropertyNode* prop) -> void
    {
        grain_ = prop ? prop->ChildAs<long long>("grain_size", 10000) : 10000;
        numSamples_ = prop ? prop->ChildAs<long long>("num_samples", 10000000L) : 10000000L;
    };
    LM_IMPL_F(Process) = [this](const Scene*, Film* film, Random* initRng, const std::function<void(Film*, Random*)>& f) -> long long
    {
        const int T = std::max(1, RefHost::numThreads);
        const long long N = numSamples_, G = grain_;
        std::atomic<long long> next(0);
        std::vector<std::unique_ptr<Random>> rngs;
        std::vector<Film::UniquePtr> films;
        for (int t = 0; t < T; t++)
        {
            rngs.emplace_back(new Random);
            rngs.back()->SetSeed(initRng->NextUInt());
            films.push_back(ComponentFactory::Clone<Film>(film));
        }
        auto work = [&](int t)
        {
            for (;;)
            {
                const long long b = next.fetch_add(G);
                if (b >= N) break;
                const long long e = std::min(N, b + G);
                for (long long i = b; i < e; i++) f(films[t].get(), rngs[t].get());
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        film->Clear();
        for (auto& fl : films) film->Accumulate(fl.get());
        film->Rescale((Float)(film->Width() * film->Height()) / N);
        return N;
    };
    LM_IMPL_F(GetNumSamples) = [this]() -> long long { return numSamples_; };
private:
    long long grain_ = 10000, numSamples_ = 0;
};

LM_COMPONENT_REGISTER_IMPL(SchedShim, "Scheduler_");

// ------------------------------------------------------------------------------------------------
// trianglemesh::mem : params {handle: <int>} selects arrays registered via RefHost::RegisterMesh

namespace {
struct MemMesh { std::vector<Float> ps, ns, ts; std::vector<unsigned int> fs; };
std::vector<std::unique_ptr<MemMesh>>& MemMeshes() { static std::vector<std::unique_ptr<MemMesh>> v; return v; }
}

auto RefHost::RegisterMesh(const float* ps, int nv, const float* ns, const float* ts, const unsigned int* fs, int nf) -> int
{
    std::unique_ptr<MemMesh> m(new MemMesh);
    m->ps.assign(ps, ps + 3 * (size_t)nv);
    if (ns) m->ns.assign(ns, ns + 3 * (size_t)nv);
    if (ts) m->ts.assign(ts, ts + 2 * (size_t)nv);
    m->fs.assign(fs, fs + 3 * (size_t)nf);
    MemMeshes().push_back(std::move(m));
    return (int)MemMeshes().size() - 1;
}

auto RefHost::ClearMeshes() -> void { MemMeshes().clear(); }

class TriangleMesh_Mem final : public TriangleMesh
{
public:
    LM_IMPL_CLASS(TriangleMesh_Mem, TriangleMesh);
    LM_IMPL_F(Load) = [this](const PropertyNode* prop, Assets*, const Primitive*) -> bool
    {
        int h = -1;
        if (!prop || !prop->ChildAs<int>("handle", h)) return false;
        if (h < 0 || h >= (int)MemMeshes().size()) return false;
        m_ = MemMeshes()[h].get();
        return true;
    };
    LM_IMPL_F(NumVertices) = [this]() -> int { return (int)(m_->ps.size() / 3); };
    LM_IMPL_F(NumFaces)    = [this]() -> int { return (int)(m_->fs.size() / 3); };
    LM_IMPL_F(Positions)   = [this]() -> const Float* { return m_->ps.data(); };
    LM_IMPL_F(Normals)     = [this]() -> const Float* { return m_->ns.empty() ? nullptr : m_->ns.data(); };
    LM_IMPL_F(Texcoords)   = [this]() -> const Float* { return m_->ts.empty() ? nullptr : m_->ts.data(); };
    LM_IMPL_F(Faces)       = [this]() -> const unsigned int* { return m_->fs.data(); };
private:
    const MemMesh* m_ = nullptr;
};

LM_COMPONENT_REGISTER_IMPL(TriangleMesh_Mem, "trianglemesh::mem");

LM_NAMESPACE_END
