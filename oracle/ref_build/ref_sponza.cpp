/* TEST INFRASTRUCTURE ONLY.  Gives the harness the reference's OWN SponzaShader (a file-static function of
 * Viewer/SponzaScene.cpp:13-104) and its file-static constants block (:11) by compiling that translation unit where it
 * lies, inside this one.  Nothing is copied: the #include below is the reference source.  The scene class around the
 * shader (model loading, camera, gamepad) is never referenced, and the library is linked with --gc-sections, so its
 * unresolved dependencies (Obj.cpp, Camera.cpp, Input.cpp) are dropped with it. */
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

/* The host's RSQRTPS, for the unit test of the device replay. */
SRB_API void srref_rsqrt(const float* in, float* out, uint64_t n)
{
	for (uint64_t i = 0; i < n; ++i)
	{
		out[i] = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(in[i])));
	}
}

} // extern "C"
