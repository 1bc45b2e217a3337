// test_host.cpp -- the C++ host layer (include/heros_host.hpp: the reference's plug-in classes over the C ABI) driven the
// way the reference's factory drives its Java classes (factory/ConstructHerosModel.java:136-146), checked against the
// CPU oracle (oracle/heros_oracle.h: test infrastructure, the checker only).
//
//   test_host cpu <golden dir>   host logic that needs no device: .properties reader, ConditionalProbability and
//                                distribution syntax, error behaviour; the engine must refuse to start without a GPU
//   test_host gpu <golden dir>   on a B200: the Javadoc worked example and boundary cases through
//                                DiseaseTransmission.infectPeople, policies, and a six-day windowed run of both
//                                transmission models whose infection / transition lists must equal the oracle's
#include "heros_host.hpp"
#include "heros_oracle.h"

#include <cstring>
#include <functional>
#include <random>

using namespace heros;

static int g_failed = 0, g_checked = 0;
#define CHECK(cond)                                                                                   \
    do {                                                                                              \
        ++g_checked;                                                                                  \
        if (!(cond)) { ++g_failed; std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); }    \
    } while (0)

template <class E, class F>
static bool throws(F&& f)
{
    try { f(); }
    catch (const E&) { return true; }
    catch (...) { return false; }
    return false;
}

// std::cerr captured while f runs (the reference reports unknown names on System.err and carries on)
template <class F>
static std::string captured_stderr(F&& f)
{
    std::ostringstream buf;
    std::streambuf* old = std::cerr.rdbuf(buf.rdbuf());
    f();
    std::cerr.rdbuf(old);
    return buf.str();
}

static bool close_rel(double a, double b, double rtol, double atol = 0.0) { return std::fabs(a - b) <= atol + rtol * std::fabs(b); }

// ---------------------------------------------------------------------------------------------------------------------
static void test_parsers(const std::string& golden)
{
    const Parameters area = loadProperties(golden + "/params_alpha_area.properties");
    CHECK(area.at("generic.diseasePropertiesModel") == "area");
    CHECK(area.at("covidT_area.t_e_mode") == "3.4" && area.at("covidT_area.calculation_threshold") == "60");
    CHECK(area.count("! the seed of the shipped scenarios") == 0 && area.at("generic.Seed") == "111");
    const Parameters dist = loadProperties(golden + "/params_alpha_distance.properties");
    CHECK(dist.at("covidT_dist.v_max") == "7.23" && dist.at("covidT_dist.psi") == "1.0");
    CHECK(throws<MedlabsException>([&] { loadProperties(golden + "/no-such-file.properties"); }));

    double t[2][128];
    parseConditionalProbability("0.46", t);
    CHECK(t[0][0] == 0.46 && t[1][127] == 0.46);
    parseConditionalProbability(area.at("covidP.FractionSymptomaticToHospitalized"), t);  // alpha-area.properties:66
    CHECK(t[0][19] == 0.02153 && t[1][20] == 0.01648 && t[0][69] == 0.44038 && t[0][100] == 0.12687 && t[0][101] == 0.0);
    parseConditionalProbability("gender{M: 0.45, F:0.5}", t);
    CHECK(t[0][30] == 0.45 && t[1][30] == 0.5);
    parseConditionalProbability("age{0-49: 0.1, 40-60: 0.9}", t);  // overlapping bands: the first match wins
    CHECK(t[0][45] == 0.1 && t[0][55] == 0.9 && t[0][61] == 0.0);
    CHECK(throws<MedlabsException>([&] { parseConditionalProbability("age{0-19 0.5}", t); }));
    CHECK(throws<MedlabsException>([&] { parseConditionalProbability("height{0-19: 0.5}", t); }));
    CHECK(throws<MedlabsException>([&] { parseConditionalProbability("abc", t); }));

    heros_duration d = parseDistribution(area.at("covidP.IncubationPeriodSymptomatic"));
    CHECK(d.kind == HEROS_DIST_TRIANGULAR && d.a == 2.5 && d.b == 3.4 && d.c == 3.8);
    d = parseDistribution(" Uniform( 1 , 2 ) ");
    CHECK(d.kind == HEROS_DIST_UNIFORM && d.a == 1.0 && d.b == 2.0);
    d = parseDistribution("Constant(3)");
    CHECK(d.kind == HEROS_DIST_CONSTANT && d.a == 3.0);
    CHECK(throws<MedlabsException>([&] { parseDistribution("Weibull(1,2)"); }));
    CHECK(throws<MedlabsException>([&] { parseDistribution("Triangular(1,2"); }));
    CHECK(throws<MedlabsException>([&] { parseDistribution("Triangular(1,2)"); }));
}

static void test_errors_without_device(const std::string& golden)
{
    Parameters bad = loadProperties(golden + "/params_alpha_area.properties");
    bad["generic.diseasePropertiesModel"] = "sir";
    CHECK(throws<MedlabsException>([&] { HerosModel m(bad); }));  // ConstructHerosModel.java:137-141
}

// there is no CPU fallback: without a usable device the model cannot be constructed
static void test_engine_refuses_to_start_without_a_gpu(const std::string& golden)
{
    const Parameters area = loadProperties(golden + "/params_alpha_area.properties");
    bool refused = false;
    try { HerosModel m(area); }
    catch (const EngineError& e) { refused = e.code == HEROS_ERR_CUDA; }
    CHECK(refused);
}

// ---------------------------------------------------------------------------------------------------------------------
// a model with one location type and `n` persons in one location of `area` m2 with one sub-location
static std::unique_ptr<HerosModel> one_room(Parameters props, int n, float area, int n_sub = 1, uint8_t infect_sub = 1)
{
    auto model = std::make_unique<HerosModel>(std::move(props), 111);
    const double corr = 1.0;
    model->check(heros_set_location_types(model->engine(), 1, &infect_sub, &corr, nullptr, nullptr, nullptr));
    const uint8_t type = 0;
    const int16_t nsub = (int16_t) n_sub;
    model->check(heros_set_locations(model->engine(), 1, &type, &nsub, &area, nullptr, nullptr));
    std::vector<int8_t> age(n, 40);
    std::vector<uint8_t> female(n, 0);
    model->check(heros_set_persons(model->engine(), n, age.data(), female.data(), nullptr, nullptr, nullptr, nullptr));
    constructDiseaseModel(*model);
    return model;
}

static void set_state(HerosModel& m, const std::vector<int32_t>& persons, const std::vector<uint8_t>& phase,
                      const std::vector<float>& exposure)
{
    m.check(heros_set_agent_state(m.engine(), (int32_t) persons.size(), persons.data(), phase.data(), exposure.data(), 0.0));
}

static void test_infect_people_known_answers(const std::string& golden)
{
    // Covid19TransmissionArea.java:112-121: one infectious person at its peak and one susceptible for one hour in 10 m2,
    // p_B = 0.51, beta = sigma_T = 1  ->  p = 1 - exp(-0.051) = 0.0497 ("0.05 (5%)")
    Parameters area = loadProperties(golden + "/params_alpha_area.properties");
    area["covidT_area.contagiousness"] = "0.51";
    const double peak = 3.4 * 24.0;
    {
        auto m = one_room(area, 2, 10.0f);
        set_state(*m, {0, 1}, {HEROS_INFECTED_SYMPTOMATIC, HEROS_SUSCEPTIBLE}, {0.0f, 0.0f});
        const auto t = m->getDiseaseTransmission();
        InfectionRecord r = t->infectPeople(0, {0, 1}, 1.0, nullptr, peak);
        CHECK(r.isCalculated() && r.exposurePhase == "Exposed" && r.location == 0);
        CHECK(close_rel(r.pInfection, 1.0 - std::exp(-0.051), 1e-12));
        CHECK(std::fabs(r.pInfection - 0.0497) < 5e-4);
        CHECK(r.infectiousPersons == std::vector<int32_t>{0});
        CHECK(close_rel(t->infectPeople(0, {0, 1}, 2.0, nullptr, peak).pInfection, 1.0 - std::exp(-0.102), 1e-12));  // 2 hours: 9.7 %
        // duration below the calculation threshold (60 s): not calculated, nothing drawn (java :138)
        r = t->infectPeople(0, {0, 1}, 59.0 / 3600.0, nullptr, peak);
        CHECK(!r.isCalculated() && r.infectedPersons.empty() && r.pInfection != r.pInfection);
        CHECK(t->infectPeople(0, {0, 1}, 1.0 / 60.0, nullptr, peak).isCalculated());  // exactly the threshold: calculated
        // not contagious yet (te <= t_e_min) and not any more (te >= t_e_max): the sum is 0, early return (java :178)
        CHECK(t->infectPeople(0, {0, 1}, 1.0, nullptr, 48.0).pInfection != t->infectPeople(0, {0, 1}, 1.0, nullptr, 48.0).pInfection);
        CHECK(t->infectPeople(0, {0, 1}, 1.0, nullptr, 9.6 * 24.0).infectedPersons.empty());
        // every ill occupant is listed as infectious, contagious or not (java :175)
        CHECK(t->infectPeople(0, {0, 1}, 1.0, nullptr, 10.0).infectiousPersons == std::vector<int32_t>{0});
        // the same evaluation by the oracle
        orc_area_props op{0.51, 1.0, 2.0, 3.4, 9.6, 60.0};
        const int32_t ids[2] = {0, 1};
        const uint8_t ph[2] = {ORC_IS, ORC_S};
        const float ex[2] = {0.f, 0.f};
        int32_t a[4], b[4], na = 0, nb = 0;
        double p = 0;
        orc_infect_people_area(&op, 111, 1, 1.0, 10.0f, 1, 2, ids, ph, ex, 2, ids, ph, ex, peak, 1.0, a, &na, b, &nb, &p);
        r = t->infectPeople(0, {0, 1}, 1.0, nullptr, peak);
        CHECK(close_rel(r.pInfection, p, 1e-12, 4.5e-16) && (int) r.infectedPersons.size() == nb);
        // setParameter: the area model has no run-time parameters -- message on stderr, no exception (java :251-255)
        const std::string msg = captured_stderr([&] { t->setParameter("psi", 2.0); });
        CHECK(msg.find("Unrecognized variable name psi") != std::string::npos);
    }
    {
        auto m = one_room(area, 2, 5.0f);  // half the area: 9.7 %
        set_state(*m, {0, 1}, {HEROS_INFECTED_SYMPTOMATIC, HEROS_SUSCEPTIBLE}, {0.0f, 0.0f});
        CHECK(std::fabs(m->getDiseaseTransmission()->infectPeople(0, {0, 1}, 1.0, nullptr, peak).pInfection - 0.0970) < 5e-4);
    }
    {
        Parameters masks = area;
        masks["covidT_area.beta"] = "0.5";  // "masks": 2.5 %
        auto m = one_room(masks, 2, 10.0f);
        set_state(*m, {0, 1}, {HEROS_INFECTED_SYMPTOMATIC, HEROS_SUSCEPTIBLE}, {0.0f, 0.0f});
        CHECK(std::fabs(m->getDiseaseTransmission()->infectPeople(0, {0, 1}, 1.0, nullptr, peak).pInfection - 0.0252) < 5e-4);
    }
    // the distance model against the oracle, before and after a DiseasePolicy (psi, mu)
    Parameters dist = loadProperties(golden + "/params_alpha_distance.properties");
    {
        auto m = one_room(dist, 6, 60.0f);
        set_state(*m, {0, 1, 2, 3, 4, 5},
                  {HEROS_INFECTED_SYMPTOMATIC, HEROS_INFECTED_ASYMPTOMATIC, HEROS_SUSCEPTIBLE, HEROS_SUSCEPTIBLE, HEROS_RECOVERED, HEROS_EXPOSED},
                  {0.0f, 20.0f, 0.0f, 0.0f, 0.0f, 70.0f});
        orc_dist_props op{2.0, 3.4, 6.2, 7.23, 4.0, 2.294, 1.0, 0.05, 0.0, 60.0};
        const int32_t ids[6] = {0, 1, 2, 3, 4, 5};
        const uint8_t ph[6] = {ORC_IS, ORC_IA, ORC_S, ORC_S, ORC_R, ORC_E};
        const float ex[6] = {0.f, 20.f, 0.f, 0.f, 0.f, 70.f};
        for (int round = 0; round < 2; ++round)
        {
            for (double now : {30.0, 70.0, 81.6, 100.0, 140.0, 230.4, 250.0})
            {
                int32_t a[8], b[8], na = 0, nb = 0;
                double p = 0;
                const int calc = orc_infect_people_dist(&op, 111, 1, 60.0f, 1, 6, ids, ph, ex, 6, ids, ph, ex, now, 2.5, a, &na, b, &nb, &p);
                const InfectionRecord r = m->getDiseaseTransmission()->infectPeople(0, {0, 1, 2, 3, 4, 5}, 2.5, nullptr, now);
                CHECK(r.isCalculated() == (calc != 0));
                CHECK((p != p && r.pInfection != r.pInfection) || close_rel(r.pInfection, p, 1e-12, 4.5e-16));
                CHECK(r.infectiousPersons == std::vector<int32_t>(a, a + na));  // only persons with v_t > 0 (java :161-166)
                CHECK(r.infectedPersons == std::vector<int32_t>(b, b + nb));
            }
            // data/thehague/policies/MaskingDistance_10-40.csv:2-3 (psi = 2.0, mu = 0.6 from day 10), fired at time 0 here
            DiseasePolicy(*m, 0.0, "covidT_dist.psi", 2.0);
            DiseasePolicy(*m, 0.0, "covidT_dist.mu", 0.6);
            op.psi = 2.0;
            op.mu = 0.6;
        }
        const std::string msg = captured_stderr([&] { DiseasePolicy(*m, 240.0, "covidT_area.beta", 0.5); });  // DiseasePolicy.java:21-25
        CHECK(msg.find("Unrecognized variable name covidT_area.beta") != std::string::npos);
        CHECK(throws<MedlabsRuntimeException>([&] { m->getDiseaseProgression()->expose(2, "Recovered"); }));
        CHECK(m->getDiseaseProgression()->getDiseasePhases().size() == 8 && m->getDiseaseProgression()->getDiseasePhases()[3] == "Infected-Symptomatic");
    }
    {
        // LocationPolicy: unknown names are reported and ignored (LocationPolicy.java:24-34)
        auto m = one_room(dist, 2, 10.0f);
        m->setLocationTypeNames({"Accommodation", "Workplace"});
        std::string msg = captured_stderr([&] { LocationPolicy p(*m, 0.0, "Casino", 0.0, 0.0, "Accommodation", "StayHome"); CHECK(!p.scheduled); });
        CHECK(msg.find("LocationPolicy: Did not recognize locationType Casino") != std::string::npos);
        msg = captured_stderr([&] { LocationPolicy p(*m, 0.0, "Workplace", 0.0, 0.0, "Moon", "StayHome"); CHECK(!p.scheduled); });
        CHECK(msg.find("LocationPolicy: Did not recognize alternativeLocationType Moon") != std::string::npos);
        LocationPolicy ok(*m, 240.0, "Workplace", 0.2, 0.5, "Accommodation", "WorkToHome");
        CHECK(ok.scheduled && ok.time == 240.0 && ok.locationType == 1 && ok.alternativeLocationType == 0);
        CHECK(throws<EngineError>([&] { ok.openCloseLocationType(); }));  // no device-side schedule has been set up
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// six simulated days of a small city, stays submitted day by day, against the oracle's event-by-event run
struct Event { int32_t person, loc; int16_t sub; double time, p; };

static void test_windowed_run_equals_the_oracle(const std::string& golden, const std::string& model_name, int trigger)
{
    Parameters props = loadProperties(golden + "/params_alpha_" + model_name + ".properties");
    if (model_name == "area") props["covidT_area.contagiousness"] = "30.0";
    else props["covidT_dist.alpha"] = "4.0";
    const int n_home = 40, n_work = 10, n_shop = 3, n_loc = n_home + n_work + n_shop, N = 640, days = 6;
    std::vector<uint8_t> infect_sub{1, 1, 1}, type(n_loc);
    std::vector<double> corr{1.0, 0.8, 1.25};
    std::vector<int16_t> nsub(n_loc);
    std::vector<float> area(n_loc);
    for (int l = 0; l < n_loc; ++l)
    {
        type[l] = l < n_home ? 0 : l < n_home + n_work ? 1 : 2;
        nsub[l] = type[l] == 0 ? 4 : type[l] == 1 ? 3 : 1;
        area[l] = (type[l] == 0 ? 100.0f : type[l] == 1 ? 30.0f : 200.0f) * nsub[l];  // ConstructHerosModel.java:381
    }
    std::mt19937_64 rng(20240807);
    auto uni = [&](double a, double b) { return a + (b - a) * ((rng() >> 11) * 0x1.0p-53); };
    std::vector<int8_t> age(N);
    std::vector<uint8_t> female(N);
    for (int p = 0; p < N; ++p) { age[p] = (int8_t) (rng() % 95); female[p] = (uint8_t) (rng() & 1); }
    std::vector<int32_t> person, loc;
    std::vector<int16_t> sub;
    std::vector<double> tin, tout;
    auto add = [&](int p, int l, int s, double a, double b) { person.push_back(p); loc.push_back(l); sub.push_back((int16_t) s); tin.push_back(a); tout.push_back(b); };
    for (int p = 0; p < N; ++p)
    {
        const int hh = p / 4, hl = hh % n_home, hs = (hh / n_home) % 4;
        const int wl = n_home + (int) (rng() % n_work), ws = (int) (rng() % 3);
        double at_home = 0.0;
        for (int d = 0; d < days; ++d)
        {
            const double leave = 24.0 * d + ((rng() % 10) == 0 ? 8.0 : uni(6.5, 9.0));  // some leave on the hour: ties
            add(p, hl, hs, at_home, leave);
            double t = leave + uni(0.2, 0.4);
            if (age[p] >= 18 && age[p] < 67)
            {
                add(p, wl, ws, t, t + 4.0);
                t += 4.0 + uni(0.05, 0.2);
                if (rng() & 1) { const double e = t + uni(0.3, 0.8); add(p, n_home + n_work + (int) (rng() % n_shop), 0, t, e); t = e + uni(0.05, 0.2); }
                add(p, wl, ws, t, t + uni(3.0, 4.0));
                t = tout.back() + uni(0.2, 0.4);
            }
            else
            {
                const double e = t + uni(0.5, 2.0);
                add(p, n_home + n_work + (int) (rng() % n_shop), 0, t, e);
                t = e + uni(0.2, 0.4);
            }
            at_home = (rng() % 8) == 0 ? std::ceil(t) : t;
        }
        add(p, hl, hs, at_home, 24.0 * days + uni(1.0, 6.0));
    }
    std::vector<int32_t> seeds;
    for (int p = 0; p < N; p += 16) seeds.push_back(p);
    const std::vector<double> zero(seeds.size(), 0.0);

    // oracle: every stay up-front, event by event
    HerosModel model(props, 4242, 0, trigger);
    constructDiseaseModel(model);  // parses covidP.* / covidT_*.* exactly as the engine will see them
    orc_sim* sim = orc_sim_create(model.modelKind() == HEROS_MODEL_AREA ? ORC_MODEL_AREA : ORC_MODEL_DIST, trigger, 4242);
    if (model.modelKind() == HEROS_MODEL_AREA)
    {
        const heros_area_params& a = static_cast<Covid19TransmissionArea&>(*model.getDiseaseTransmission()).params;
        orc_area_props op{a.contagiousness, a.beta, a.t_e_min, a.t_e_mode, a.t_e_max, a.calculation_threshold};
        orc_sim_set_area(sim, &op);
    }
    else
    {
        const heros_dist_params& a = static_cast<Covid19TransmissionDistance&>(*model.getDiseaseTransmission()).params;
        orc_dist_props op{a.L, a.I, a.C, a.v_max, a.v_0, a.r, a.psi, a.alpha, a.mu, a.calculation_threshold};
        orc_sim_set_dist(sim, &op);
    }
    {
        auto op = std::make_unique<orc_prog_props>();
        const heros_prog_params& pp = model.getDiseaseProgression()->params();
        std::memcpy(op->fraction, pp.fraction, sizeof op->fraction);
        for (int k = 0; k < ORC_NPERIOD; ++k) op->period[k] = orc_duration{pp.period[k].kind, pp.period[k].a, pp.period[k].b, pp.period[k].c};
        orc_sim_set_prog(sim, op.get());
    }
    orc_sim_set_location_types(sim, 3, infect_sub.data(), corr.data());
    orc_sim_set_locations(sim, n_loc, type.data(), nsub.data(), area.data());
    orc_sim_set_persons(sim, N, age.data(), female.data());
    orc_sim_expose(sim, (int) seeds.size(), seeds.data(), zero.data());
    orc_sim_add_visits(sim, (int64_t) person.size(), person.data(), loc.data(), sub.data(), tin.data(), tout.data());

    // engine: the same tables through the C ABI, stays submitted before the day they begin in
    model.check(heros_set_location_types(model.engine(), 3, infect_sub.data(), corr.data(), nullptr, nullptr, nullptr));
    model.check(heros_set_locations(model.engine(), n_loc, type.data(), nsub.data(), area.data(), nullptr, nullptr));
    model.check(heros_set_persons(model.engine(), N, age.data(), female.data(), nullptr, nullptr, nullptr, nullptr));
    model.getDiseaseProgression()->expose(seeds, Covid19Progression::exposed, 0.0);  // ConstructHerosModel.java:1123
    for (int d = 0; d < days; ++d)
    {
        std::vector<int32_t> dp, dl;
        std::vector<int16_t> ds;
        std::vector<double> da, db;
        for (size_t i = 0; i < person.size(); ++i)
            if (tin[i] >= 24.0 * d && tin[i] < 24.0 * (d + 1))
            { dp.push_back(person[i]); dl.push_back(loc[i]); ds.push_back(sub[i]); da.push_back(tin[i]); db.push_back(tout[i]); }
        model.check(heros_submit_visits(model.engine(), (int64_t) dp.size(), dp.data(), dl.data(), ds.data(), da.data(), db.data()));
        heros_step_stats st{};
        model.check(heros_step(model.engine(), 24.0 * (d + 1), &st));
        CHECK(st.windows == 1 && st.gpu_launches > 0);
        orc_sim_run(sim, 24.0 * (d + 1));
        int64_t gc[8], oc[8];
        model.check(heros_get_phase_counts(model.engine(), gc));
        orc_sim_get_phase_counts(sim, oc);
        CHECK(std::memcmp(gc, oc, sizeof gc) == 0);
        CHECK(model.getSimulatorTime() == 24.0 * (d + 1));
    }
    // infection event lists
    int64_t n = 0;
    model.check(heros_get_infections(model.engine(), 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, &n));
    std::vector<Event> g((size_t) n), o((size_t) orc_sim_n_infections(sim));
    CHECK(g.size() == o.size() && g.size() > 20);
    {
        std::vector<int32_t> a(std::max<size_t>(g.size(), o.size()) + 1), b(a.size());
        std::vector<int16_t> c(a.size());
        std::vector<double> t(a.size()), p(a.size());
        if (n) model.check(heros_get_infections(model.engine(), 0, n, a.data(), b.data(), c.data(), t.data(), p.data(), &n));
        for (size_t i = 0; i < g.size(); ++i) g[i] = Event{a[i], b[i], c[i], t[i], p[i]};
        orc_sim_get_infections(sim, a.data(), b.data(), c.data(), t.data(), p.data());
        for (size_t i = 0; i < o.size(); ++i) o[i] = Event{a[i], b[i], c[i], t[i], p[i]};
    }
    auto by_time = [](const Event& x, const Event& y) { return x.time != y.time ? x.time < y.time : x.person < y.person; };
    std::sort(g.begin(), g.end(), by_time);
    std::sort(o.begin(), o.end(), by_time);
    bool same = g.size() == o.size();
    for (size_t i = 0; same && i < g.size(); ++i)
        same = g[i].person == o[i].person && g[i].loc == o[i].loc && g[i].sub == o[i].sub && g[i].time == o[i].time &&
               close_rel(g[i].p, o[i].p, 1e-12, 4.5e-16);  // north-star tolerance on p (+ 4 ulp(1): 1 - exp(x) cancels)
    CHECK(same);
    // disease-phase trajectories
    int64_t nt = 0;
    model.check(heros_get_transitions(model.engine(), 0, 0, nullptr, nullptr, nullptr, &nt));
    CHECK(nt == orc_sim_n_transitions(sim) && nt > (int64_t) seeds.size());
    if (nt == orc_sim_n_transitions(sim) && nt > 0)
    {
        std::vector<int32_t> gp((size_t) nt), op2((size_t) nt);
        std::vector<uint8_t> gph((size_t) nt), oph((size_t) nt);
        std::vector<double> gt((size_t) nt), ot((size_t) nt);
        model.check(heros_get_transitions(model.engine(), 0, nt, gp.data(), gph.data(), gt.data(), &nt));
        orc_sim_get_transitions(sim, op2.data(), oph.data(), ot.data());
        std::vector<size_t> gi((size_t) nt), oi((size_t) nt);
        for (size_t i = 0; i < (size_t) nt; ++i) gi[i] = oi[i] = i;
        std::sort(gi.begin(), gi.end(), [&](size_t x, size_t y) { return gt[x] != gt[y] ? gt[x] < gt[y] : gp[x] < gp[y]; });
        std::sort(oi.begin(), oi.end(), [&](size_t x, size_t y) { return ot[x] != ot[y] ? ot[x] < ot[y] : op2[x] < op2[y]; });
        bool eq = true;
        for (size_t i = 0; eq && i < (size_t) nt; ++i) eq = gp[gi[i]] == op2[oi[i]] && gph[gi[i]] == oph[oi[i]] && gt[gi[i]] == ot[oi[i]];
        CHECK(eq);
    }
    std::printf("  %s model, trigger %d: %zu infections, %lld transitions identical to the oracle\n", model_name.c_str(), trigger,
                g.size(), (long long) nt);
    orc_sim_destroy(sim);
}

// the device-side schedule through the C ABI on hand-made tables: accepted when consistent, refused (before any kernel
// runs) when a table points outside the location table
static void test_device_schedule_tables(const std::string& golden)
{
    HerosModel model(loadProperties(golden + "/params_alpha_area.properties"), 5);
    constructDiseaseModel(model);
    const uint8_t infect_sub[3] = {1, 1, 1};
    const double corr[3] = {1.0, 1.0, 1.0};
    model.check(heros_set_location_types(model.engine(), 3, infect_sub, corr, nullptr, nullptr, nullptr));
    const uint8_t type[3] = {0, 1, 2};
    const int16_t nsub[3] = {2, 1, 1};
    const float area[3] = {200.0f, 30.0f, 200.0f};
    model.check(heros_set_locations(model.engine(), 3, type, nsub, area, nullptr, nullptr));
    const int8_t age[4] = {30, 31, 32, 33};
    const uint8_t female[4] = {0, 1, 0, 1}, role[4] = {7, 7, 7, 7};
    const int32_t home_loc[4] = {0, 0, 0, 0}, work_loc[4] = {1, 1, 1, -1};
    const int16_t home_sub[4] = {0, 0, 1, 1}, work_sub[4] = {0, 0, 0, 0};
    model.check(heros_set_persons(model.engine(), 4, age, female, role, home_loc, home_sub, work_loc));
    heros_activity acts[3] = {};
    acts[0].kind = 2; acts[0].locator = 0; acts[0].choice = -1; acts[0].until = 8.0;                       // at home until 08:00
    acts[1].kind = 0; acts[1].locator = 1; acts[1].choice = -1; acts[1].min = acts[1].mode = acts[1].max = 8.0;  // 8 h at work
    acts[2].kind = 2; acts[2].locator = 0; acts[2].choice = -1; acts[2].until = 24.0;                      // home until midnight
    std::vector<int32_t> pat_start(512, 0), pat_count(512, 0);
    for (int ph = 0; ph < 8; ++ph)
        for (int d = 0; d < 4; ++d) pat_count[(ph * 16 + 7) * 4 + d] = 3;
    const int32_t choice_start[1] = {0}, choice_type[1] = {0};
    const double choice_cum[1] = {1.0};
    std::vector<int32_t> nearest(9, -1), n_cand(9, 0), cand(9, -1);
    auto set = [&](const int32_t* near) {
        return heros_set_schedule(model.engine(), 3, acts, pat_start.data(), pat_count.data(), 0, choice_start, choice_type, choice_cum,
                                  3, 0, 1, near, n_cand.data(), cand.data(), work_sub, 5);
    };
    std::vector<int32_t> wrong = nearest;
    wrong[4] = 3;  // a location that does not exist
    CHECK(set(wrong.data()) == HEROS_ERR_INVALID);
    CHECK(std::string(heros_last_error(model.engine())).find("nearest-location table") != std::string::npos);
    model.check(set(nearest.data()));
    model.check(heros_advance_schedule(model.engine(), 0));
    heros_step_stats st{};
    model.check(heros_step(model.engine(), 24.0, &st));
    CHECK(st.windows == 1 && st.visits >= 7);  // three commuters: home, work, home; one person at home all day
    model.check(heros_advance_schedule(model.engine(), 1));
    model.check(heros_step(model.engine(), 48.0, &st));
    CHECK(model.getSimulatorTime() == 48.0);
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::printf("usage: %s cpu|gpu <golden dir>\n", argv[0]); return 2; }
    const std::string mode = argv[1], golden = argv[2];
    test_parsers(golden);
    test_errors_without_device(golden);
    if (mode == "cpu")
        test_engine_refuses_to_start_without_a_gpu(golden);
    else if (mode == "schedule")
        test_device_schedule_tables(golden);
    else
    {
        test_device_schedule_tables(golden);
        test_infect_people_known_answers(golden);
        test_windowed_run_equals_the_oracle(golden, "area", HEROS_TRIGGER_EXIT_ONLY);
        test_windowed_run_equals_the_oracle(golden, "distance", HEROS_TRIGGER_EXIT_ONLY);
        test_windowed_run_equals_the_oracle(golden, "distance", HEROS_TRIGGER_ENTER_AND_EXIT);
    }
    std::printf("%s: %d checks, %d failed\n", mode.c_str(), g_checked, g_failed);
    return g_failed ? 1 : 0;
}
