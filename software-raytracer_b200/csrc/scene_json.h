// scene_json.h - host-side scene model and the reference's Scenes/*.json format.
//
// Mirrors Scene (Scene.hpp:12-119) and Object::ToJSON (Object.hpp:27-43,143-147,218-222):
// same schema, same defaults, same "partial load on a bad entry" behaviour - but with an
// error code and message instead of a silent return. Self-contained JSON reader/writer
// (the reference vendors nlohmann/json 3.11.2; nothing of it is copied here).
#pragma once
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "mesh.h"

namespace rtb {

struct HostScene {
    std::string file_name;                 // Scene::fileName
    std::string scene_name;                // Scene::sceneName  ("SceneName")
    std::vector<rt_object> objects;        // Scene::sceneObjects, JSON order
    std::vector<std::string> names;        // Object::name      ("Name")
    std::vector<HostMesh> meshes;          // per object; empty unless type == RT_OBJ_MESH (extension, mesh.h)

    // Scene::Load (Scene.hpp:27-80). Returns RT_OK, RT_ERR_IO (file missing: scene left empty,
    // like the reference's silent return) or RT_ERR_PARSE (objects parsed before the bad entry
    // are kept, like the reference's catch block). `err` receives the message.
    int Load(const std::string& path, std::string& err);
    int LoadFromString(const std::string& text, std::string& err);
    // Scene::SaveAs / Save (Scene.hpp:88-104): dump(4) formatting, keys in byte order.
    int SaveAs(const std::string& path, std::string& err);
    std::string Dump() const;
    void Unload() { objects.clear(); names.clear(); meshes.clear(); }                       // Scene.hpp:81-87
    void AddObject(const rt_object& o, const std::string& name = "") {      // Scene.hpp:105-107
        objects.push_back(o); names.push_back(name); meshes.resize(objects.size());
    }
    bool RemoveObject(size_t index) {                                        // Scene.hpp:108-115
        if (index >= objects.size()) return false;
        objects.erase(objects.begin() + index); names.erase(names.begin() + index);
        if (index < meshes.size()) meshes.erase(meshes.begin() + index);
        return true;
    }
};

// Material() defaults (Common.hpp:313-318) on a zeroed object of the given type.
rt_object make_object(int type);

}  // namespace rtb
