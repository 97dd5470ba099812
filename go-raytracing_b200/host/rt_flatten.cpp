// rt_flatten.cpp — pointer scene graph -> SoA rtx_scene_desc (include/rtx_b200.h).
//
// The Go drop-in does this inside package rt with a type switch over the concrete types (it needs the
// unexported fields Quad.u/v, Triangle.v0..v2, BVHNode.left/right, Lambertian.tex, Volume.boundary).
// An unknown Hittable / Material / Texture is a flatten-time error — there is no CPU fallback.
#include <unordered_map>

#include "rt.hpp"

namespace rt {

namespace {
struct Flattener {
    FlatScene& fs;
    std::unordered_map<const Texture*, int> texIds;
    std::unordered_map<const Material*, int> matIds;
    std::unordered_map<const Hittable*, int> sphIds, quadIds, triIds, planeIds, circleIds, groupIds;
    std::unordered_map<const Perlin*, int> perlinIds;
    explicit Flattener(FlatScene& f) : fs(f) {}

    static void push3(std::vector<double>& v, const Vec3& a) { v.push_back(a.X); v.push_back(a.Y); v.push_back(a.Z); }

    int texture(const TexturePtr& t) {
        if (!t) throw std::runtime_error("flatten: nil texture");
        auto it = texIds.find(t.get());
        if (it != texIds.end()) return it->second;
        int even = -1, odd = -1, type;
        Color c{0, 0, 0};
        double inv = 0;
        if (auto s = dynamic_cast<const SolidColor*>(t.get())) { type = RTX_TEX_SOLID; c = s->Albedo; }
        else if (auto ch = dynamic_cast<const CheckerTexture*>(t.get())) {
            type = RTX_TEX_CHECKER; inv = ch->invScale; even = texture(ch->even); odd = texture(ch->odd);
        } else if (auto nt = dynamic_cast<const NoiseTexture*>(t.get())) {   // scale travels in tex_inv_scale (not inverted), the table index in tex_even
            type = RTX_TEX_NOISE; inv = nt->scale;
            auto pit = perlinIds.find(nt->noise.get());
            if (pit == perlinIds.end()) {
                even = (int)(fs.perlin_perm.size() / 768);
                for (int k = 0; k < 256; k++) push3(fs.perlin_vec, nt->noise->randvec[k]);
                for (int k = 0; k < 256; k++) fs.perlin_perm.push_back(nt->noise->permX[k]);
                for (int k = 0; k < 256; k++) fs.perlin_perm.push_back(nt->noise->permY[k]);
                for (int k = 0; k < 256; k++) fs.perlin_perm.push_back(nt->noise->permZ[k]);
                perlinIds[nt->noise.get()] = even;
            } else even = pit->second;
        } else if (auto it2 = dynamic_cast<const ImageTexture*>(t.get())) {
            if (!it2->image || it2->image->Height() <= 0)
                throw std::runtime_error("flatten: ImageTexture without image data (the reference would draw its cyan debug colour; load the image first)");
            type = RTX_TEX_IMAGE;
            even = (int)fs.image_width.size();
            fs.image_width.push_back(it2->image->imageWidth); fs.image_height.push_back(it2->image->imageHeight);
            fs.image_offset.push_back((int64_t)(fs.image_rgb.size() / 3));
            fs.image_rgb.insert(fs.image_rgb.end(), it2->image->data.begin(), it2->image->data.end());
        } else throw std::runtime_error("flatten: unsupported Texture type (SolidColor, CheckerTexture, NoiseTexture and ImageTexture are on the device path)");
        int id = (int)fs.tex_type.size();
        fs.tex_type.push_back(type); push3(fs.tex_color, c); fs.tex_inv_scale.push_back(inv);
        fs.tex_even.push_back(even); fs.tex_odd.push_back(odd);
        texIds[t.get()] = id;
        return id;
    }
    int material(const MaterialPtr& m) {
        if (!m) throw std::runtime_error("flatten: nil material");
        auto it = matIds.find(m.get());
        if (it != matIds.end()) return it->second;
        int type, tex = -1;
        Color albedo{0, 0, 0};
        double fuzz = 0, ior = 0;
        if (auto l = dynamic_cast<const Lambertian*>(m.get())) { type = RTX_MAT_LAMBERTIAN; tex = texture(l->tex); }
        else if (auto me = dynamic_cast<const Metal*>(m.get())) { type = RTX_MAT_METAL; albedo = me->Albedo; fuzz = me->Fuzz; }
        else if (auto d = dynamic_cast<const Dielectric*>(m.get())) { type = RTX_MAT_DIELECTRIC; ior = d->RefractionIndex; }
        else if (auto dl = dynamic_cast<const DiffuseLight*>(m.get())) { type = RTX_MAT_DIFFUSE_LIGHT; tex = texture(dl->tex); }
        else if (auto is = dynamic_cast<const Isotropic*>(m.get())) { type = RTX_MAT_ISOTROPIC; tex = texture(is->tex); }
        else throw std::runtime_error("flatten: unsupported Material type");
        int id = (int)fs.mat_type.size();
        fs.mat_type.push_back(type); fs.mat_tex.push_back(tex); push3(fs.mat_albedo, albedo);
        fs.mat_fuzz.push_back(fuzz); fs.mat_ior.push_back(ior);
        matIds[m.get()] = id;
        return id;
    }
    // Returns (kind, index) of a primitive, adding it on first sight (shared objects are stored once).
    bool primitive(const Hittable* h, int& kind, int& index) {
        if (auto s = dynamic_cast<const Sphere*>(h)) {
            kind = RTX_GEOM_SPHERE;
            auto it = sphIds.find(h);
            if (it != sphIds.end()) { index = it->second; return true; }
            index = (int)fs.sph_mat.size();
            push3(fs.sph_center, s->Center0); push3(fs.sph_velocity, s->Velocity);
            fs.sph_radius.push_back(s->rawRadius); fs.sph_mat.push_back(material(s->Mat));
            sphIds[h] = index;
            return true;
        }
        if (auto q = dynamic_cast<const Quad*>(h)) {
            kind = RTX_GEOM_QUAD;
            auto it = quadIds.find(h);
            if (it != quadIds.end()) { index = it->second; return true; }
            index = (int)fs.quad_mat.size();
            push3(fs.quad_q, q->Q); push3(fs.quad_u, q->u); push3(fs.quad_v, q->v); fs.quad_mat.push_back(material(q->mat));
            quadIds[h] = index;
            return true;
        }
        if (auto t = dynamic_cast<const Triangle*>(h)) {
            kind = RTX_GEOM_TRIANGLE;
            auto it = triIds.find(h);
            if (it != triIds.end()) { index = it->second; return true; }
            index = (int)fs.tri_mat.size();
            push3(fs.tri_v0, t->v0); push3(fs.tri_v1, t->v1); push3(fs.tri_v2, t->v2);
            fs.tri_mat.push_back(material(t->mat)); fs.tri_rank.push_back(0);
            triIds[h] = index;
            return true;
        }
        if (auto c = dynamic_cast<const Circle*>(h)) {
            kind = RTX_GEOM_CIRCLE;
            auto it = circleIds.find(h);
            if (it != circleIds.end()) { index = it->second; return true; }
            index = (int)fs.circle_mat.size();
            push3(fs.circle_center, c->center); push3(fs.circle_normal, c->normal); fs.circle_radius.push_back(c->radius);
            fs.circle_mat.push_back(material(c->mat));
            circleIds[h] = index;
            return true;
        }
        if (auto p = dynamic_cast<const Plane*>(h)) {
            kind = RTX_GEOM_PLANE;
            auto it = planeIds.find(h);
            if (it != planeIds.end()) { index = it->second; return true; }
            index = (int)fs.plane_mat.size();
            push3(fs.plane_point, p->Point); push3(fs.plane_normal, p->Normal); fs.plane_mat.push_back(material(p->Mat));
            planeIds[h] = index;
            return true;
        }
        return false;
    }
    // DFS of a reference-order BVH: objects in the order BVHNode.Hit tests them (rt/bvh.go:219-239); each
    // leaf is referenced from both sides (rt/bvh.go:141) and counted once.
    static void dfs(const Hittable* h, std::vector<const Hittable*>& out) {
        if (!h) return;
        if (auto n = dynamic_cast<const BVHNode*>(h)) {
            if (n->deferred) { out.push_back(h); return; }   // a mesh root whose tree was never built: one object of the enclosing tree
            if (n->left && n->left == n->right) { dfs(n->left.get(), out); return; }
            dfs(n->left.get(), out);
            dfs(n->right.get(), out);
        } else if (auto l = dynamic_cast<const BVHLeaf*>(h)) {
            for (auto& o : l->objects) out.push_back(o.get());
        } else out.push_back(h);
    }
    int listGroup(const HittableList* l) {
        auto it = groupIds.find(l);
        if (it != groupIds.end()) return it->second;
        int begin = (int)fs.list_item_kind.size();
        for (auto& o : l->Objects) {
            int k, idx;
            if (!primitive(o.get(), k, idx))
                throw std::runtime_error("flatten: a nested HittableList may only hold primitives (Box = 6 quads, rt/primitives.go:5)");
            fs.list_item_kind.push_back(k); fs.list_item_index.push_back(idx);
        }
        int id = (int)fs.group_kind.size();
        fs.group_kind.push_back(RTX_GEOM_LIST); fs.group_begin.push_back(begin); fs.group_count.push_back((int)l->Objects.size());
        groupIds[l] = id;
        return id;
    }
    int meshGroup(const BVHNode* root) {
        auto it = groupIds.find(root);
        if (it != groupIds.end()) return it->second;
        if (root->soup && root->deferred) {   // LoadOBJ's flat face arrays: copied as they are; the library derives the test order on the device
            const BVHNode::Soup& sp = *root->soup;
            int begin = (int)fs.tri_mat.size();
            fs.tri_v0.insert(fs.tri_v0.end(), sp.v0.begin(), sp.v0.end());
            fs.tri_v1.insert(fs.tri_v1.end(), sp.v1.begin(), sp.v1.end());
            fs.tri_v2.insert(fs.tri_v2.end(), sp.v2.begin(), sp.v2.end());
            fs.tri_mat.insert(fs.tri_mat.end(), sp.n, material(sp.mat));
            fs.tri_rank.insert(fs.tri_rank.end(), sp.n, 0);
            fs.have_tri_rank = false;
            int id = (int)fs.group_kind.size();
            fs.group_kind.push_back(RTX_GEOM_MESH); fs.group_begin.push_back(begin); fs.group_count.push_back((int)sp.n);
            groupIds[root] = id;
            return id;
        }
        if (root->src.empty()) throw std::runtime_error("flatten: mesh BVH without its source slice (build it with NewBVHNode / LoadOBJ)");
        int begin = (int)fs.tri_mat.size();
        std::unordered_map<const Hittable*, int> local;
        if (!root->deferred) local.reserve(root->src.size() * 2);
        fs.tri_v0.reserve(fs.tri_v0.size() + 3 * root->src.size()); fs.tri_v1.reserve(fs.tri_v1.size() + 3 * root->src.size()); fs.tri_v2.reserve(fs.tri_v2.size() + 3 * root->src.size());
        for (auto& o : root->src) {  // face order; mesh triangles are never shared with other geometry
            auto t = dynamic_cast<const Triangle*>(o.get());
            if (!t) throw std::runtime_error("flatten: a nested BVH must be a triangle mesh (rt/obj_loader.go:109)");
            if (!root->deferred) local[o.get()] = (int)fs.tri_mat.size() - begin;
            push3(fs.tri_v0, t->v0); push3(fs.tri_v1, t->v1); push3(fs.tri_v2, t->v2);
            fs.tri_mat.push_back(material(t->mat)); fs.tri_rank.push_back(0);
        }
        if (root->deferred) fs.have_tri_rank = false;   // no tree on this side: the library derives the canonical test order on the device
        else {
            std::vector<const Hittable*> order;
            order.reserve(root->src.size());
            dfs(root, order);
            for (size_t r = 0; r < order.size(); r++) fs.tri_rank[begin + local.at(order[r])] = (int)r;
        }
        int id = (int)fs.group_kind.size();
        fs.group_kind.push_back(RTX_GEOM_MESH); fs.group_begin.push_back(begin); fs.group_count.push_back((int)root->src.size());
        groupIds[root] = id;
        return id;
    }
    void entry(const Hittable* h) {
        int volume = -1;
        if (auto v = dynamic_cast<const Volume*>(h)) {
            volume = (int)fs.vol_mat.size();
            fs.vol_neg_inv_density.push_back(v->negInvDensity); fs.vol_mat.push_back(material(v->phaseFunction));
            h = v->boundary.get();
        }
        int xfBegin = (int)fs.xf_type.size(), xfCount = 0;
        while (true) {
            if (auto t = dynamic_cast<const Translate*>(h)) {
                fs.xf_type.push_back(RTX_XF_TRANSLATE); push3(fs.xf_a, t->Offset); push3(fs.xf_b, {0, 0, 0}); h = t->Obj.get();
            } else if (auto r = dynamic_cast<const RotateY*>(h)) {
                fs.xf_type.push_back(RTX_XF_ROTATE_Y); push3(fs.xf_a, {r->SinTheta, r->CosTheta, 0}); push3(fs.xf_b, {0, 0, 0}); h = r->Obj.get();
            } else if (auto s = dynamic_cast<const Scale*>(h)) {
                fs.xf_type.push_back(RTX_XF_SCALE); push3(fs.xf_a, s->Factor); push3(fs.xf_b, s->InvFactor); h = s->Obj.get();
            } else break;
            xfCount++;
        }
        int kind, index;
        if (primitive(h, kind, index)) {
        } else if (auto l = dynamic_cast<const HittableList*>(h)) { kind = RTX_GEOM_LIST; index = listGroup(l); }
        else if (auto b = dynamic_cast<const BVHNode*>(h)) { kind = RTX_GEOM_MESH; index = meshGroup(b); }
        else if (dynamic_cast<const Volume*>(h)) throw std::runtime_error("flatten: a Volume inside a transform is outside the device path");
        else throw std::runtime_error("flatten: unsupported Hittable type (user-defined hittables cannot run on the device)");
        fs.entry_geom_kind.push_back(kind); fs.entry_geom_index.push_back(index);
        fs.entry_xf_begin.push_back(xfBegin); fs.entry_xf_count.push_back(xfCount);
        fs.entry_volume.push_back(volume); fs.entry_rank.push_back((int)fs.entry_rank.size());
    }
};
}  // namespace

std::shared_ptr<FlatScene> Flatten(const HittablePtr& world, const Camera& camera) {
    auto fs = std::make_shared<FlatScene>();
    Flattener fl(*fs);
    if (auto list = dynamic_cast<const HittableList*>(world.get())) {
        fs->world_is_bvh = false;
        for (auto& o : list->Objects) fl.entry(o.get());
    } else if (auto bvh = dynamic_cast<const BVHNode*>(world.get())) {
        fs->world_is_bvh = true;
        for (auto& o : bvh->src) fl.entry(o.get());
        // test-order ranks from the tree (an object added twice gets consecutive ranks in insertion order)
        std::vector<const Hittable*> order;
        Flattener::dfs(bvh, order);
        std::unordered_map<const Hittable*, std::vector<int>> where;
        for (size_t i = 0; i < bvh->src.size(); i++) where[bvh->src[i].get()].push_back((int)i);
        std::unordered_map<const Hittable*, size_t> used;
        for (size_t r = 0; r < order.size(); r++) {
            auto& v = where[order[r]];
            size_t& u = used[order[r]];
            if (u < v.size()) fs->entry_rank[v[u++]] = (int)r;
        }
    } else {
        throw std::runtime_error("flatten: world must be a *HittableList or the *BVHNode returned by NewBVHNodeFromList");
    }
    for (auto& l : camera.Lights) {  // Camera.Lights order matters: uniform pick by index (rt/camera.go:502-505)
        int k, idx;
        if (dynamic_cast<const Quad*>(l.get()) && fl.primitive(l.get(), k, idx)) fs->light_quad.push_back(idx);
        else fs->light_quad.push_back(-1);
    }
    if (camera.Environment && camera.Environment->IsValid()) fs->env = camera.Environment;
    return fs;
}

rtx_scene_desc FlatScene::Desc() const {
    rtx_scene_desc d{};
    d.abi_version = RTX_ABI_VERSION;
    d.world_is_bvh = world_is_bvh;
    d.n_textures = (int)tex_type.size(); d.tex_type = tex_type.data(); d.tex_color = tex_color.data();
    d.tex_inv_scale = tex_inv_scale.data(); d.tex_even = tex_even.data(); d.tex_odd = tex_odd.data();
    d.n_materials = (int)mat_type.size(); d.mat_type = mat_type.data(); d.mat_tex = mat_tex.data();
    d.mat_albedo = mat_albedo.data(); d.mat_fuzz = mat_fuzz.data(); d.mat_ior = mat_ior.data();
    d.n_spheres = (int)sph_mat.size(); d.sph_center = sph_center.data(); d.sph_velocity = sph_velocity.data();
    d.sph_radius = sph_radius.data(); d.sph_mat = sph_mat.data();
    d.n_quads = (int)quad_mat.size(); d.quad_q = quad_q.data(); d.quad_u = quad_u.data(); d.quad_v = quad_v.data(); d.quad_mat = quad_mat.data();
    d.n_tris = (int)tri_mat.size(); d.tri_v0 = tri_v0.data(); d.tri_v1 = tri_v1.data(); d.tri_v2 = tri_v2.data();
    d.tri_mat = tri_mat.data(); d.tri_rank = have_tri_rank ? tri_rank.data() : nullptr;
    d.n_planes = (int)plane_mat.size(); d.plane_point = plane_point.data(); d.plane_normal = plane_normal.data(); d.plane_mat = plane_mat.data();
    d.n_circles = (int)circle_mat.size(); d.circle_center = circle_center.data(); d.circle_normal = circle_normal.data();
    d.circle_radius = circle_radius.data(); d.circle_mat = circle_mat.data();
    d.n_perlin = (int)(perlin_perm.size() / 768); d.perlin_vec = perlin_vec.data(); d.perlin_perm = perlin_perm.data();
    d.n_images = (int)image_width.size(); d.image_width = image_width.data(); d.image_height = image_height.data();
    d.image_offset = image_offset.data(); d.image_rgb = image_rgb.data();
    d.n_groups = (int)group_kind.size(); d.group_kind = group_kind.data(); d.group_begin = group_begin.data(); d.group_count = group_count.data();
    d.n_list_items = (int)list_item_kind.size(); d.list_item_kind = list_item_kind.data(); d.list_item_index = list_item_index.data();
    d.n_xforms = (int)xf_type.size(); d.xf_type = xf_type.data(); d.xf_a = xf_a.data(); d.xf_b = xf_b.data();
    d.n_volumes = (int)vol_mat.size(); d.vol_neg_inv_density = vol_neg_inv_density.data(); d.vol_mat = vol_mat.data();
    d.n_entries = (int)entry_geom_kind.size(); d.entry_geom_kind = entry_geom_kind.data(); d.entry_geom_index = entry_geom_index.data();
    d.entry_xf_begin = entry_xf_begin.data(); d.entry_xf_count = entry_xf_count.data(); d.entry_volume = entry_volume.data();
    d.entry_rank = entry_rank.data();
    d.n_lights = (int)light_quad.size(); d.light_quad = light_quad.data();
    if (env) {
        d.env_width = env->width; d.env_height = env->height; d.env_rgb = env->rgb.data();
        d.env_rotation = env->rotation; d.env_importance_sampling = env->useImportanceSampling;
    }
    return d;
}

}  // namespace rt
