/* TEST INFRASTRUCTURE ONLY.  Gives the harness the reference's OWN SponzaShader (a file-static function of
 * Viewer/SponzaScene.cpp:13-104) and its file-static constants block (:11) by compiling that translation unit where it
 * lies, inside this one.  Nothing is copied: the #include below is the reference source.  The scene class around the
 * shader is used only by srref_sponza_scene_* below (its model loader is ref_obj.cpp's Obj.cpp; the camera controller gets
 * link-time stand-ins at the end of this file). */
#include "Viewer/SponzaScene.cpp"

#include "../../include/softrast_b200.h"

sr::PixelShaderFn* srref_sponza_shader_fn() { return &sr::SponzaShader; }

extern "C"
{

/* Fills the reference's g_constants the way SponzaScene::Init / Update do (SponzaScene.cpp:121-187) from the POD the
 * C ABI uses.  sun_dir / ambient are the broadcast values of the three __m256 lanes. */
SRB_API void srref_set_sponza_constants(const srb_sponza_constants* k)
{
	for (int i = 0; i < 3; ++i)
	{
		sr::g_constants.m_sunDir[i] = _mm256_set1_ps(k->sun_dir[i]);
		sr::g_constants.m_ambCol[i] = _mm256_set1_ps(k->ambient[i]);
	}
	static_assert(SRB_SPONZA_POINT_LIGHTS == sr::SponzaScene::Constants::c_numPointLights, "light count");
	for (uint32_t i = 0; i < SRB_SPONZA_POINT_LIGHTS; ++i)
	{
		sr::SponzaScene::PointLight& l = sr::g_constants.m_pointLights[i];
		l.m_pos = kt::Vec3(k->lights[i].pos[0], k->lights[i].pos[1], k->lights[i].pos[2]);
		l.m_colour = kt::Vec3(k->lights[i].colour[0], k->lights[i].colour[1], k->lights[i].colour[2]);
		l.m_intensity = k->lights[i].intensity;
		l.m_falloff = k->lights[i].falloff;
	}
}

/* SponzaScene::Init / Update themselves (SponzaScene.cpp:110-215), for the test of srb_sponza_scene_*: the scene is
 * constructed on a model path that does not exist (Load fails, no meshes, so Update issues no draws), Init seeds the
 * lights, every Update(dt) animates them; the constants block is exported after each call. */
SRB_API void* srref_sponza_scene_create(void)
{
	sr::SponzaScene* sc = new sr::SponzaScene("/nonexistent/srref_no_model.obj", 0);
	sc->Init(720, 1280);
	return sc;
}

SRB_API void srref_sponza_scene_destroy(void* scene) { delete static_cast<sr::SponzaScene*>(scene); }

SRB_API void srref_sponza_scene_update(void* scene, void* renderContext, void* frameBuffer, float dt)
{
	static_cast<sr::SponzaScene*>(scene)->Update(*static_cast<sr::RenderContext*>(renderContext),
	                                             *static_cast<sr::FrameBuffer*>(frameBuffer), dt);
}

SRB_API void srref_get_sponza_constants(srb_sponza_constants* k)
{
	memset(k, 0, sizeof(*k));
	for (int i = 0; i < 3; ++i)
	{
		k->sun_dir[i] = _mm256_cvtss_f32(sr::g_constants.m_sunDir[i]);
		k->ambient[i] = _mm256_cvtss_f32(sr::g_constants.m_ambCol[i]);
	}
	for (uint32_t i = 0; i < SRB_SPONZA_POINT_LIGHTS; ++i)
	{
		sr::SponzaScene::PointLight const& l = sr::g_constants.m_pointLights[i];
		k->lights[i].pos[0] = l.m_pos.x, k->lights[i].pos[1] = l.m_pos.y, k->lights[i].pos[2] = l.m_pos.z;
		k->lights[i].colour[0] = l.m_colour.x, k->lights[i].colour[1] = l.m_colour.y, k->lights[i].colour[2] = l.m_colour.z;
		k->lights[i].intensity = l.m_intensity;
		k->lights[i].falloff = l.m_falloff;
	}
}

/* The host's RSQRTPS, for the unit test of the device replay. */
SRB_API void srref_rsqrt(const float* in, float* out, uint64_t n)
{
	for (uint64_t i = 0; i < n; ++i)
	{
		out[i] = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(in[i])));
	}
}

} // extern "C"

/* Link-time stand-ins for the camera controller the scene class holds (Viewer/Camera.cpp needs the Windows gamepad input
 * layer): the lights do not depend on the camera, and with no meshes Update never asks for the view-projection. */
namespace sr
{
void FreeCamController::SetPos(kt::Vec3 const&) {}
void FreeCamController::UpdateViewGamepad(float const) {}
void FreeCamController::SetProjectionParams(ProjectionParams const& _params) { m_projectionParams = _params; }
Camera& FreeCamController::GetCam() { return m_camera; }
kt::Mat4 const& Camera::GetCachedViewProj() const { return m_cachedWorldToClip; }
}
